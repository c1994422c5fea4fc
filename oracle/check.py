"""The checker used by tests / smoke(): the unmodified reference (oracle/_ref) when it is
available, else our CPU restatement (oracle/restate).  TEST INFRASTRUCTURE ONLY."""
from . import refbind


def _use_ref():
    return refbind.available()


def kind() -> str:
    return "reference" if _use_ref() else "port"


def encode(data, dict_size, level, read_size=0, flt_mask=7, with_props=True):
    if _use_ref():
        return refbind.encode(data, dict_size, level, read_size, flt_mask, with_props)
    from . import restate
    return restate.encode(data, dict_size, level, read_size, flt_mask, with_props)


def decode(stream, raw_size_hint):
    if _use_ref():
        return refbind.decode(stream, raw_size_hint)
    from . import restate
    return restate.decode(stream, raw_size_hint)
