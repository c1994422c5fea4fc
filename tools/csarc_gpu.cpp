// csarc_gpu -- a csarc-compatible archiver that keeps thousands of tasks in flight on the GPU.
//
// The reference archiver (archiver/csarc.cpp) drives libcsc with at most 8 worker threads, one handle per task
// (csarc.cpp:338-409, -t clamp :200-201), and splits a single file into at most -p pieces of >= 1 MiB
// (:532-543).  One handle is one serially dependent stream, so a B200 needs ~10^3 tasks in flight.  This tool
// writes and reads the SAME archive format -- 24-byte header, one libcsc stream (10-byte props + blocks) per task
// stored as extents of the archive file, packed + compressed index at the end (csarc.cpp:219-286,
// csa_indexpack.cpp:160-209, csa_typedef.h:12-55) -- so its archives are listed, tested and unpacked by the stock
// tool and it unpacks the stock tool's.  What differs is the schedule: tasks are cut small (default 1 MiB, any
// number of files may be cut: the index allows up to 127 fragments per file), gathered into batches of
// gigabytes and handed to the library's batched calls, which also return csarc's per-fragment Adler-32
// (csa_adler32.cpp, taken with initial value 0: csa_io.h:250,325-342) from the pass the bytes make through
// HBM.  Host side only: plain C++ above the C ABI of include/csc_b200.h + csc_enc.h / csc_dec.h.
//
// usage: csarc_gpu a [-m1..5] [-d<n>[k|m]] [-c<chunk>[k|m]] [-b<batch MiB>] [-f] archive path...
//        csarc_gpu x [-o<dir>] archive        csarc_gpu t archive        csarc_gpu l archive
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <dirent.h>
#include <sys/stat.h>
#include <utime.h>
#include <unistd.h>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include <csc_enc.h>
#include <csc_dec.h>
#include <csc_b200.h>

using std::string;
using std::vector;

namespace {

// ---- index structures (csa_typedef.h:12-55) and their packed form (csa_indexpack.cpp:66-158)
struct Frag { uint32_t bid, checksum; uint64_t posblock, size, posfile; };
struct Entry { int64_t edate = 0, esize = 0, eattr = 0; vector<Frag> frags; };
struct Extent { uint64_t off, size; };
typedef std::map<string, Entry> FileIndex;
typedef std::map<uint64_t, vector<Extent>> BlockIndex;

void put(vector<char> &b, uint64_t v, int n) { for (int i = 0; i < n; i++) { b.push_back((char)(v & 0xFF)); v >>= 8; } }
uint64_t get(const char *&p, int n) { uint64_t v = 0; for (int i = n - 1; i >= 0; i--) v = (v << 8) | (uint8_t)p[i]; p += n; return v; }

vector<char> pack_index(const FileIndex &fi, const BlockIndex &bi) {
    vector<char> b;
    put(b, fi.size(), 4);
    for (auto &kv : fi) {
        put(b, kv.first.size(), 4);
        b.insert(b.end(), kv.first.begin(), kv.first.end());
        put(b, (uint64_t)kv.second.edate, 8); put(b, (uint64_t)kv.second.esize, 8); put(b, (uint64_t)kv.second.eattr, 8);
        b.push_back((char)kv.second.frags.size());
        for (auto &f : kv.second.frags) { put(b, f.bid, 4); put(b, f.checksum, 4); put(b, f.posblock, 8); put(b, f.size, 8); put(b, f.posfile, 8); }
    }
    put(b, bi.size(), 4);
    for (auto &kv : bi) {
        put(b, kv.first, 8);
        put(b, kv.second.size(), 4);
        for (auto &e : kv.second) { put(b, e.off, 8); put(b, e.size, 8); }
    }
    return b;
}
bool unpack_index(const vector<char> &b, FileIndex &fi, BlockIndex &bi) {
    const char *p = b.data(), *end = b.data() + b.size();
    if (b.size() < 8) return false;
    uint32_t n = (uint32_t)get(p, 4);
    for (uint32_t i = 0; i < n; i++) {
        if (p + 4 > end) return false;
        uint32_t l = (uint32_t)get(p, 4);
        if (p + l + 25 > end) return false;
        string name(p, l); p += l;
        Entry e;
        e.edate = (int64_t)get(p, 8); e.esize = (int64_t)get(p, 8); e.eattr = (int64_t)get(p, 8);
        int nf = (int8_t)*p++;
        for (int k = 0; k < nf; k++) {
            if (p + 32 > end) return false;
            Frag f; f.bid = (uint32_t)get(p, 4); f.checksum = (uint32_t)get(p, 4); f.posblock = get(p, 8); f.size = get(p, 8); f.posfile = get(p, 8);
            e.frags.push_back(f);
        }
        fi[name] = e;
    }
    if (p + 4 > end) return false;
    n = (uint32_t)get(p, 4);
    for (uint32_t i = 0; i < n; i++) {
        if (p + 12 > end) return false;
        uint64_t id = get(p, 8);
        uint32_t nb = (uint32_t)get(p, 4);
        vector<Extent> v;
        for (uint32_t k = 0; k < nb; k++) { if (p + 16 > end) return false; Extent e; e.off = get(p, 8); e.size = get(p, 8); v.push_back(e); }
        bi[id] = v;
    }
    return true;
}

// csarc's checksum is adler32() started at 0 (csa_io.h:325); the library's by-product starts at 1 like zlib.
// A(0) = A(1) - 1 and B(0) = B(1) - n (mod 65521).
uint32_t adler_from_one_to_zero(uint32_t a1, uint64_t n) {
    const uint32_t M = 65521;
    uint32_t a = a1 & 0xFFFF, b = a1 >> 16;
    a = (a + M - 1) % M;
    b = (uint32_t)((b + M - n % M) % M);
    return a | (b << 16);
}

// ---- memory streams for the handle API (index stream)
struct MemIn { ISeqInStream is; const char *p; size_t n, at; };
struct MemOut { ISeqOutStream os; vector<char> *v; };
int mem_read(void *o, void *buf, size_t *size) { MemIn *m = (MemIn *)o; size_t k = std::min(*size, m->n - m->at); memcpy(buf, m->p + m->at, k); m->at += k; *size = k; return 0; }
size_t mem_write(void *o, const void *buf, size_t size) { MemOut *m = (MemOut *)o; m->v->insert(m->v->end(), (const char *)buf, (const char *)buf + size); return size; }

// The index stream is written through the handle API.  A zero-filling allocator makes the one byte per flush segment that
// libcsc's Coder::Flush skips instead of storing (it keeps whatever the buffer held) deterministic on the reference library;
// the GPU library always behaves like that.
void *zero_alloc(void *, size_t size) { return calloc(1, size ? size : 1); }
void zero_free(void *, void *p) { free(p); }
ISzAlloc g_zero_alloc = {zero_alloc, zero_free};

int64_t decimal_time(time_t t) {   // csa_common.cpp: YYYYMMDDHHMMSS (UTC)
    struct tm *g = gmtime(&t);
    if (!g) return 0;
    return (g->tm_year + 1900LL) * 10000000000LL + (g->tm_mon + 1) * 100000000LL + g->tm_mday * 1000000LL + g->tm_hour * 10000 + g->tm_min * 100 + g->tm_sec;
}
time_t unix_time(int64_t d) {
    if (d <= 0) return 0;
    struct tm t; memset(&t, 0, sizeof(t));
    t.tm_year = (int)(d / 10000000000LL) - 1900; t.tm_mon = (int)(d / 100000000LL % 100) - 1; t.tm_mday = (int)(d / 1000000 % 100);
    t.tm_hour = (int)(d / 10000 % 100); t.tm_min = (int)(d / 100 % 100); t.tm_sec = (int)(d % 100);
    return timegm(&t);
}

void scan(const string &path0, FileIndex &fi) {   // csarc.cpp: scandir (regular files and directories, recursive)
    string path = path0;
    while (path.size() > 1 && path.back() == '/') path.pop_back();
    struct stat sb;
    if (lstat(path.c_str(), &sb)) { perror(path.c_str()); return; }
    if (S_ISREG(sb.st_mode)) {
        Entry e; e.edate = decimal_time(sb.st_mtime); e.esize = sb.st_size; e.eattr = 'u' + ((int64_t)sb.st_mode << 8);
        fi[path] = e;
    } else if (S_ISDIR(sb.st_mode)) {
        Entry e; e.edate = decimal_time(sb.st_mtime); e.esize = 0; e.eattr = 'u' + ((int64_t)sb.st_mode << 8);
        fi[path == "/" ? path : path + "/"] = e;
        DIR *d = opendir(path.c_str());
        if (!d) { perror(path.c_str()); return; }
        for (dirent *dp = readdir(d); dp; dp = readdir(d))
            if (strcmp(dp->d_name, ".") && strcmp(dp->d_name, "..")) scan((path == "/" ? path : path + "/") + dp->d_name, fi);
        closedir(d);
    }
}
string ext_of(const string &n) {
    size_t dot = n.find_last_of('.'), slash = n.find_last_of('/');
    if (dot == string::npos || (slash != string::npos && dot < slash)) return "";
    string e = n.substr(dot + 1, 4);
    for (auto &c : e) c = (char)tolower(c);
    return e;
}
void make_dirs(const string &file) {
    for (size_t i = 1; i < file.size(); i++)
        if (file[i] == '/') { string d = file.substr(0, i); mkdir(d.c_str(), 0777); }
}
uint64_t parse_size(const char *s) {
    char *e; uint64_t v = strtoull(s, &e, 10);
    if (*e == 'k' || *e == 'K') v <<= 10; else if (*e == 'm' || *e == 'M') v <<= 20; else if (*e == 'g' || *e == 'G') v <<= 30;
    return v;
}

struct Piece { string file; uint64_t posfile, size, posblock; };   // a fragment of a file inside a task
struct Task { vector<Piece> pieces; uint64_t size = 0; };

// ---- add
int cmd_add(const string &arc, const vector<string> &paths, int level, uint64_t dict, uint64_t chunk, uint64_t batch, bool force) {
    if (!force && access(arc.c_str(), F_OK) == 0) { fprintf(stderr, "Archive %s already exists, use -f to force overwrite\n", arc.c_str()); return 1; }
    FileIndex fi;
    for (auto &p : paths) scan(p, fi);
    // files in extension order (csarc.cpp:496-513 groups like with like), then cut into tasks
    vector<FileIndex::iterator> files;
    for (auto it = fi.begin(); it != fi.end(); ++it) if (!it->first.empty() && it->first.back() != '/' && it->second.esize > 0) files.push_back(it);
    std::stable_sort(files.begin(), files.end(), [](const FileIndex::iterator &a, const FileIndex::iterator &b) { return ext_of(a->first) < ext_of(b->first); });
    vector<Task> tasks;
    Task cur;
    string cur_ext;
    for (auto it : files) {
        const uint64_t sz = (uint64_t)it->second.esize;
        if (sz >= chunk) {   // a file of its own: up to 127 pieces (the index keeps the fragment count in one signed byte)
            if (cur.size) { tasks.push_back(cur); cur = Task(); }
            const uint64_t piece = std::max<uint64_t>(chunk, (sz + 126) / 127);
            for (uint64_t off = 0; off < sz; off += piece) {
                Task t; t.size = std::min(piece, sz - off);
                t.pieces.push_back({it->first, off, t.size, 0});
                tasks.push_back(t);
            }
        } else {             // small files of one kind share a task (solid), up to the chunk size
            if (cur.size && (cur.size + sz > chunk || (ext_of(it->first) != cur_ext && cur.size > 64 * 1024))) { tasks.push_back(cur); cur = Task(); }
            cur_ext = ext_of(it->first);
            cur.pieces.push_back({it->first, 0, sz, cur.size});
            cur.size += sz;
        }
    }
    if (cur.size) tasks.push_back(cur);

    FILE *out = fopen(arc.c_str(), "wb");
    if (!out) { perror(arc.c_str()); return 1; }
    char hdr[24]; memset(hdr, 0, sizeof(hdr));
    fwrite(hdr, 1, 24, out);
    uint64_t arc_pos = 24, raw_total = 0;
    BlockIndex bi;
    // batches of tasks: read, one batched encode (+ Adler-32 of every task), append the streams in task order
    for (size_t t0 = 0; t0 < tasks.size();) {
        size_t t1 = t0; uint64_t bytes = 0;
        while (t1 < tasks.size() && (t1 == t0 || bytes + tasks[t1].size <= batch)) bytes += tasks[t1++].size;
        vector<uint8_t> raw(bytes);
        vector<const uint8_t *> in; vector<size_t> in_size, cap, out_size; vector<uint8_t *> outp;
        vector<const uint8_t *> fr_ptr; vector<size_t> fr_size; vector<std::pair<string, size_t>> fr_dst;   // fragments of multi-file tasks
        uint64_t at = 0, out_total = 0;
        for (size_t t = t0; t < t1; t++) {
            in.push_back(raw.data() + at); in_size.push_back(tasks[t].size);
            for (auto &p : tasks[t].pieces) {
                FILE *f = fopen(p.file.c_str(), "rb");
                size_t got = 0;
                if (f) { fseeko(f, (off_t)p.posfile, SEEK_SET); got = fread(raw.data() + at + p.posblock, 1, p.size, f); fclose(f); }
                if (got != p.size) { fprintf(stderr, "%s: short read\n", p.file.c_str()); fclose(out); return 1; }
            }
            at += tasks[t].size;
            cap.push_back(CSCB200_EncodeBound(tasks[t].size)); out_total += cap.back();
        }
        vector<uint8_t> comp(out_total);
        uint64_t o = 0;
        for (size_t k = 0; k < cap.size(); k++) { outp.push_back(comp.data() + o); o += cap[k]; }
        out_size.resize(in.size());
        vector<uint32_t> adler(in.size());
        int r = CSCB200_EncodeChunksAdler((int)in.size(), in.data(), in_size.data(), (uint32_t)dict, level, 7, 1, outp.data(), cap.data(), out_size.data(), adler.data());
        if (r < 0) { fprintf(stderr, "CSCB200_EncodeChunksAdler failed: %d\n", r); fclose(out); return 1; }
        for (size_t t = t0; t < t1; t++) {
            const size_t k = t - t0;
            fwrite(outp[k], 1, out_size[k], out);
            bi[t] = {Extent{arc_pos, out_size[k]}};
            arc_pos += out_size[k];
            for (auto &p : tasks[t].pieces) {
                Frag f; f.bid = (uint32_t)t; f.posblock = p.posblock; f.size = p.size; f.posfile = p.posfile;
                f.checksum = adler_from_one_to_zero(adler[k], tasks[t].size);   // right for single-piece tasks; the others follow
                fi[p.file].frags.push_back(f);
                if (tasks[t].pieces.size() > 1) { fr_ptr.push_back(in[k] + p.posblock); fr_size.push_back(p.size); fr_dst.push_back({p.file, fi[p.file].frags.size() - 1}); }
            }
        }
        if (!fr_ptr.empty()) {   // solid tasks: one checksum per member file
            vector<uint32_t> zero(fr_ptr.size(), 0), sums(fr_ptr.size());
            r = CSCB200_Adler32((int)fr_ptr.size(), fr_ptr.data(), fr_size.data(), zero.data(), sums.data());
            if (r < 0) { fprintf(stderr, "CSCB200_Adler32 failed: %d\n", r); fclose(out); return 1; }
            for (size_t k = 0; k < fr_dst.size(); k++) fi[fr_dst[k].first].frags[fr_dst[k].second].checksum = sums[k];
        }
        raw_total += bytes;
        t0 = t1;
    }
    // index: packed, compressed as one -m2 stream with a 256 KiB dictionary, at the end (csarc.cpp:219-286)
    vector<char> idx = pack_index(fi, bi), cidx;
    {
        CSCProps p;
        CSCEncProps_Init(&p, 256 * 1024, 2);
        MemIn mi; mi.is.Read = mem_read; mi.p = idx.data(); mi.n = idx.size(); mi.at = 0;
        MemOut mo; mo.os.Write = mem_write; mo.v = &cidx;
        uint8_t props[CSC_PROP_SIZE];
        CSCEnc_WriteProperties(&p, props, 0);
        cidx.insert(cidx.end(), (char *)props, (char *)props + CSC_PROP_SIZE);
        CSCEncHandle h = CSCEnc_Create(&p, (ISeqOutStream *)&mo, &g_zero_alloc);
        if (!h || CSCEnc_Encode(h, (ISeqInStream *)&mi, NULL) < 0 || CSCEnc_Encode_Flush(h) < 0) { fprintf(stderr, "index compression failed\n"); fclose(out); return 1; }
        CSCEnc_Destroy(h);
    }
    fwrite(cidx.data(), 1, cidx.size(), out);
    vector<char> h8;
    h8.push_back('C'); h8.push_back('S'); h8.push_back('A'); put(h8, 0x20130331u, 4); h8.push_back('1');
    put(h8, arc_pos, 8); put(h8, cidx.size(), 4); put(h8, idx.size(), 4);
    fseeko(out, 0, SEEK_SET);
    fwrite(h8.data(), 1, 24, out);
    fclose(out);
    printf("%zu files, %zu tasks, %llu -> %llu bytes\n", files.size(), tasks.size(), (unsigned long long)raw_total, (unsigned long long)(arc_pos + cidx.size()));
    return 0;
}

// ---- list / test / extract
int read_index(const string &arc, FILE *&f, FileIndex &fi, BlockIndex &bi) {
    f = fopen(arc.c_str(), "rb");
    if (!f) { perror(arc.c_str()); return -1; }
    char hdr[24];
    if (fread(hdr, 1, 24, f) != 24 || hdr[0] != 'C' || hdr[1] != 'S' || hdr[2] != 'A' || hdr[7] != '1') { fprintf(stderr, "Invalid csarc file\n"); return -1; }
    const char *p = hdr + 3;
    if (get(p, 4) != 0x20130331u) { fprintf(stderr, "Invalid csarc file\n"); return -1; }
    p = hdr + 8;
    uint64_t pos = get(p, 8); uint32_t csize = (uint32_t)get(p, 4), rsize = (uint32_t)get(p, 4);
    vector<char> c(csize);
    fseeko(f, (off_t)pos, SEEK_SET);
    if (csize < CSC_PROP_SIZE || fread(c.data(), 1, csize, f) != csize) { fprintf(stderr, "truncated index\n"); return -1; }
    CSCProps pr;
    CSCDec_ReadProperties(&pr, (uint8_t *)c.data());
    MemIn mi; mi.is.Read = mem_read; mi.p = c.data(); mi.n = c.size(); mi.at = CSC_PROP_SIZE;
    vector<char> raw; raw.reserve(rsize);
    MemOut mo; mo.os.Write = mem_write; mo.v = &raw;
    CSCDecHandle h = CSCDec_Create(&pr, (ISeqInStream *)&mi, NULL);
    if (!h || CSCDec_Decode(h, (ISeqOutStream *)&mo, NULL) < 0 || raw.size() != rsize) { fprintf(stderr, "index corrupted\n"); return -1; }
    CSCDec_Destroy(h);
    return unpack_index(raw, fi, bi) ? 0 : -1;
}

int cmd_extract(const string &arc, const string &to_dir, bool write, uint64_t batch) {
    FILE *f; FileIndex fi; BlockIndex bi;
    if (read_index(arc, f, fi, bi) < 0) return 1;
    struct Out { const string *name; const Frag *fr; };
    std::map<uint64_t, vector<Out>> by_task;   // task id -> its fragments
    for (auto &kv : fi) {
        const string dst = (to_dir.empty() ? string("") : to_dir + "/") + kv.first;
        if (write) {
            make_dirs(dst);
            if (!kv.first.empty() && kv.first.back() == '/') mkdir(dst.c_str(), 0777);
            else { FILE *o = fopen(dst.c_str(), "wb"); if (o) fclose(o); }
        }
        for (auto &fr : kv.second.frags) if (fr.size) by_task[fr.bid].push_back({&kv.first, &fr});
    }
    vector<uint64_t> ids;
    for (auto &kv : by_task) ids.push_back(kv.first);
    int bad = 0;
    for (size_t t0 = 0; t0 < ids.size();) {
        // a batch of tasks: their streams gathered from the archive's extents, one batched decode (+ Adler-32)
        size_t t1 = t0; uint64_t bytes = 0;
        vector<uint64_t> raw_size;
        while (t1 < ids.size()) {
            uint64_t rs = 0;
            for (auto &o : by_task[ids[t1]]) rs = std::max(rs, o.fr->posblock + o.fr->size);
            if (t1 > t0 && bytes + rs > batch) break;
            raw_size.push_back(rs); bytes += rs; t1++;
        }
        vector<vector<uint8_t>> comp(t1 - t0);
        vector<const uint8_t *> in; vector<size_t> in_size, cap, out_size(t1 - t0);
        vector<uint8_t> raw(bytes + 64 * (t1 - t0));
        vector<uint8_t *> outp;
        uint64_t at = 0;
        for (size_t t = t0; t < t1; t++) {
            auto &c = comp[t - t0];
            for (auto &e : bi[ids[t]]) { size_t k = c.size(); c.resize(k + e.size); fseeko(f, (off_t)e.off, SEEK_SET); if (fread(c.data() + k, 1, e.size, f) != e.size) { fprintf(stderr, "truncated archive\n"); return 1; } }
            in.push_back(c.data()); in_size.push_back(c.size());
            outp.push_back(raw.data() + at); cap.push_back(raw_size[t - t0]); at += raw_size[t - t0] + 64;
        }
        vector<uint32_t> adler(t1 - t0);
        int r = CSCB200_DecodeChunksAdler((int)in.size(), in.data(), in_size.data(), outp.data(), cap.data(), out_size.data(), adler.data());
        if (r < 0) { fprintf(stderr, "Extraction error, archive corrupted (%d)\n", r); return 1; }
        vector<const uint8_t *> fr_ptr; vector<size_t> fr_size; vector<Out> fr_out;
        for (size_t t = t0; t < t1; t++) {
            auto &outs = by_task[ids[t]];
            const size_t k = t - t0;
            for (auto &o : outs) {
                if (o.fr->posblock + o.fr->size > out_size[k]) { fprintf(stderr, "******** %s extraction/verify failed (short task)\n", o.name->c_str()); bad++; continue; }
                if (outs.size() == 1 && o.fr->posblock == 0 && o.fr->size == out_size[k]) {
                    if (adler_from_one_to_zero(adler[k], out_size[k]) != o.fr->checksum) { fprintf(stderr, "******** %s extraction/verify failed\n", o.name->c_str()); bad++; }
                } else { fr_ptr.push_back(outp[k] + o.fr->posblock); fr_size.push_back(o.fr->size); fr_out.push_back(o); }
                if (write) {
                    const string dst = (to_dir.empty() ? string("") : to_dir + "/") + *o.name;
                    FILE *w = fopen(dst.c_str(), "r+b");
                    if (!w) { perror(dst.c_str()); bad++; continue; }
                    fseeko(w, (off_t)o.fr->posfile, SEEK_SET);
                    fwrite(outp[k] + o.fr->posblock, 1, o.fr->size, w);
                    fclose(w);
                }
            }
        }
        if (!fr_ptr.empty()) {
            vector<uint32_t> zero(fr_ptr.size(), 0), sums(fr_ptr.size());
            r = CSCB200_Adler32((int)fr_ptr.size(), fr_ptr.data(), fr_size.data(), zero.data(), sums.data());
            if (r < 0) { fprintf(stderr, "CSCB200_Adler32 failed: %d\n", r); return 1; }
            for (size_t k = 0; k < fr_out.size(); k++)
                if (sums[k] != fr_out[k].fr->checksum) { fprintf(stderr, "******** %s extraction/verify failed\n", fr_out[k].name->c_str()); bad++; }
        }
        t0 = t1;
    }
    if (write)
        for (auto &kv : fi) {   // dates and modes as csarc restores them (csa_file.cpp)
            const string dst = (to_dir.empty() ? string("") : to_dir + "/") + kv.first;
            if ((kv.second.eattr & 0xFF) == 'u') chmod(dst.c_str(), (mode_t)((kv.second.eattr >> 8) & 07777));
            struct utimbuf ub; ub.actime = time(NULL); ub.modtime = unix_time(kv.second.edate);
            if (ub.modtime > 0) utime(dst.c_str(), &ub);
        }
    fclose(f);
    printf("%zu tasks %s\n", ids.size(), bad ? "FAILED" : "OK");
    return bad ? 1 : 0;
}

int cmd_list(const string &arc) {
    FILE *f; FileIndex fi; BlockIndex bi;
    if (read_index(arc, f, fi, bi) < 0) return 1;
    fclose(f);
    for (auto &kv : fi) printf("%s %lld\n", kv.first.c_str(), (long long)kv.second.esize);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: csarc_gpu a|x|t|l [-m#] [-d#[k|m]] [-c#[k|m]] [-b#] [-f] [-o<dir>] archive [path...]\n"); return 2; }
    const string cmd = argv[1];
    int level = 2; uint64_t dict = 32000000, chunk = 1 << 20, batch = 2048ull << 20; bool force = false; string to_dir;
    vector<string> rest;
    for (int i = 2; i < argc; i++) {
        const char *a = argv[i];
        if (a[0] == '-' && a[1] == 'm') level = atoi(a + 2);
        else if (a[0] == '-' && a[1] == 'd') dict = parse_size(a + 2);
        else if (a[0] == '-' && a[1] == 'c') chunk = std::max<uint64_t>(parse_size(a + 2), 65536);
        else if (a[0] == '-' && a[1] == 'b') batch = parse_size(a + 2) << 20;
        else if (a[0] == '-' && a[1] == 'f') force = true;
        else if (a[0] == '-' && a[1] == 'o') to_dir = a + 2;
        else rest.push_back(a);
    }
    if (rest.empty()) { fprintf(stderr, "no archive name\n"); return 2; }
    if (!CSCB200_Available() && cmd != "l") fprintf(stderr, "csarc_gpu: no CUDA device (the library has no CPU fallback)\n");
    const string arc = rest[0];
    rest.erase(rest.begin());
    if (cmd == "a") return cmd_add(arc, rest, level, dict, chunk, batch, force);
    if (cmd == "x") return cmd_extract(arc, to_dir, true, batch);
    if (cmd == "t") return cmd_extract(arc, "", false, batch);
    if (cmd == "l") return cmd_list(arc);
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}
