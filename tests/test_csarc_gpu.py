"""tools/csarc_gpu.cpp -- the GPU-aware, csarc-compatible archiver.

CPU part: the tool linked against oracle/b200_refshim.cpp (the batched calls over the unmodified reference library)
must write archives the stock csarc lists, tests and unpacks, and must unpack the stock csarc's -- that pins the
archive format logic (header, task extents, packed index, fragment checksums) without a GPU.
GPU part: the same tool on the product library writes the SAME BYTES as on the reference library (every task
stream is byte-identical to libcsc's), with hundreds of tasks in one batched call.
"""
import os
import subprocess

import numpy as np
import pytest

from csc_b200 import synth


def _tools(ref, *names):
    try:
        return [ref.cli(n) for n in names]
    except FileNotFoundError:
        pytest.skip("oracle/_ref archiver tools not built and /root/reference absent")


def _tree(root, scale=1):
    files = {"big.txt": synth.text(3_500_000 * scale, 1), "prog.exe": synth.exe(700_000 * scale, 2),
             "sub/data.bin": synth.mixed(1_300_000 * scale, 3), "sub/empty.dat": np.zeros(0, np.uint8),
             "sub/deep/table.dat": synth.delta(150_000, 4, 5)}
    for i in range(6):
        files[f"sub/n{i}.txt"] = synth.text(20_000 + 1000 * i, 10 + i)
    for n, d in files.items():
        os.makedirs(os.path.dirname(os.path.join(root, "src", n)), exist_ok=True)
        d.tofile(os.path.join(root, "src", n))
    return files


def _run(cwd, *a):
    r = subprocess.run(list(a), cwd=cwd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (a, r.stdout[-400:], r.stderr[-400:])
    assert "verify failed" not in r.stderr and "FAILED" not in r.stdout, (a, r.stderr[-400:])
    return r


def _same_tree(root, out, files):
    for n, d in files.items():
        got = np.fromfile(os.path.join(root, out, "src", n), dtype=np.uint8)
        assert got.size == d.size and bool((got == d).all()), (out, n)


def test_archive_format_against_stock_csarc(ref, tmp_path):
    tool, stock = _tools(ref, "csarc_gpu_refshim", "csarc_ref")
    files = _tree(tmp_path)
    # ours -> stock: list, test, extract (files cut into many tasks, small files solid by extension, empty file, directories)
    _run(tmp_path, tool, "a", "-m2", "-c512k", "-f", "ours.csa", "src")
    listing = _run(tmp_path, stock, "l", "ours.csa").stdout
    for n, d in files.items():
        assert f"src/{n} {d.size}" in listing
    _run(tmp_path, stock, "t", "ours.csa")
    os.mkdir(tmp_path / "o1")
    _run(tmp_path, stock, "x", "-o", "o1/", "ours.csa")
    _same_tree(tmp_path, "o1", files)
    # ours -> ours
    _run(tmp_path, tool, "t", "ours.csa")
    _run(tmp_path, tool, "x", "-oo2", "ours.csa")
    _same_tree(tmp_path, "o2", files)
    # stock -> ours (solid tasks, blocks of one task interleaved with the others' in the archive file)
    _run(tmp_path, stock, "a", "-r", "-m1", "-t4", "-f", "stock.csa", "src")
    _run(tmp_path, tool, "t", "stock.csa")
    _run(tmp_path, tool, "x", "-oo3", "stock.csa")
    _same_tree(tmp_path, "o3", files)
    # a damaged task stream is reported, not unpacked silently
    b = bytearray((tmp_path / "ours.csa").read_bytes())
    b[5000] ^= 0x40
    (tmp_path / "bad.csa").write_bytes(bytes(b))
    r = subprocess.run([tool, "t", "bad.csa"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0


def test_large_file_is_cut_into_at_most_127_fragments(ref, tmp_path):
    """the packed index keeps a file's fragment count in one signed byte (csa_indexpack.cpp:86,107): a file larger than
    127 chunks gets larger pieces instead of more of them"""
    tool, stock = _tools(ref, "csarc_gpu_refshim", "csarc_ref")
    d = synth.mixed(20 << 20, 31)
    os.makedirs(tmp_path / "src")
    d.tofile(tmp_path / "src" / "big.bin")
    out = _run(tmp_path, tool, "a", "-m1", "-c64k", "-f", "big.csa", "src").stdout
    assert "127 tasks" in out, out
    listing = _run(tmp_path, stock, "l", "-v", "big.csa").stdout
    assert "Fragment 126" in listing and "Fragment 127" not in listing
    _run(tmp_path, stock, "t", "big.csa")
    os.mkdir(tmp_path / "o")
    _run(tmp_path, stock, "x", "-o", "o/", "big.csa")
    got = np.fromfile(tmp_path / "o" / "src" / "big.bin", dtype=np.uint8)
    assert got.size == d.size and bool((got == d).all())


@pytest.mark.gpu
def test_csarc_gpu_on_the_gpu(gpu_lib, ref, tmp_path):
    tool, shim, stock = _tools(ref, "csarc_gpu", "csarc_gpu_refshim", "csarc_ref")
    files = _tree(tmp_path, scale=6)          # ~35 MB, ~150 tasks of 256 KiB in ONE batched call
    _run(tmp_path, tool, "a", "-m1", "-c256k", "-f", "gpu.csa", "src")
    _run(tmp_path, shim, "a", "-m1", "-c256k", "-f", "cpu.csa", "src")
    a, b = np.fromfile(tmp_path / "gpu.csa", dtype=np.uint8), np.fromfile(tmp_path / "cpu.csa", dtype=np.uint8)
    if not (a.size == b.size and bool((a == b).all())):
        n = min(a.size, b.size)
        diff = np.flatnonzero(a[:n] != b[:n])
        lg = _run(tmp_path, stock, "l", "-v", "gpu.csa").stdout.splitlines()
        lc = _run(tmp_path, stock, "l", "-v", "cpu.csa").stdout.splitlines()
        bad = [(x, y) for x, y in zip(lg, lc) if x != y][:6]
        raise AssertionError(f"archive written on the GPU differs from the one written with libcsc: sizes {a.size} / {b.size}, "
                             f"{diff.size} differing bytes, first at {diff[:8].tolist()}, listing differences {bad}")
    _run(tmp_path, stock, "t", "gpu.csa")
    _run(tmp_path, tool, "t", "gpu.csa")
    _run(tmp_path, tool, "x", "-oo1", "gpu.csa")
    _same_tree(tmp_path, "o1", files)
    # the stock archiver's archive through the GPU decoder, -m3
    _run(tmp_path, stock, "a", "-r", "-m3", "-t8", "-f", "stock.csa", "src")
    _run(tmp_path, tool, "x", "-oo2", "stock.csa")
    _same_tree(tmp_path, "o2", files)
