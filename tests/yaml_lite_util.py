"""ctypes access to ndt_b200/libyaml_lite.so and the same event listings produced with
libyaml 0.2.5 itself (PyYAML's C extension embeds it): the reference the YAML layer is
pinned against, since the reference (scene.c:573-2175) delegates all YAML syntax to libyaml."""
import ctypes as C
import os

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ndt_b200", "libyaml_lite.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ylite_events_from_yaml.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.ylite_yaml_from_events.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.ylite_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def lite_events(text: bytes):
    """(error code, listing) from yaml_lite's parser"""
    out = C.c_void_p(); n = C.c_size_t()
    rc = lib().ylite_events_from_yaml(text, len(text), C.byref(out), C.byref(n))
    s = C.string_at(out, n.value)
    lib().ylite_free(out)
    return rc, s.decode("utf-8", "surrogateescape")


def lite_emit(listing: str, width=0):
    """(error code, text) from yaml_lite's emitter"""
    b = listing.encode("utf-8", "surrogateescape")
    out = C.c_void_p(); n = C.c_size_t()
    rc = lib().ylite_yaml_from_events(b, len(b), width, C.byref(out), C.byref(n))
    s = C.string_at(out, n.value) if out.value else b""
    lib().ylite_free(out)
    return rc, s


def _esc(v):
    return v.replace("\\", "\\\\").replace("\n", "\\n").replace("\r", "\\r").replace("\t", "\\t").replace("\0", "\\0")


def libyaml_events(text: bytes):
    """the same listing from libyaml's parser; (error, listing-so-far)"""
    out = []
    err = 0
    try:
        for ev in yaml.parse(text, Loader=yaml.CLoader):
            if isinstance(ev, yaml.StreamStartEvent): out.append("+STR")
            elif isinstance(ev, yaml.StreamEndEvent): out.append("-STR")
            elif isinstance(ev, yaml.DocumentStartEvent): out.append("+DOC ---" if ev.explicit else "+DOC")
            elif isinstance(ev, yaml.DocumentEndEvent): out.append("-DOC ..." if ev.explicit else "-DOC")
            elif isinstance(ev, yaml.MappingStartEvent): out.append("+MAP {}" if ev.flow_style else "+MAP")
            elif isinstance(ev, yaml.MappingEndEvent): out.append("-MAP")
            elif isinstance(ev, yaml.SequenceStartEvent): out.append("+SEQ []" if ev.flow_style else "+SEQ")
            elif isinstance(ev, yaml.SequenceEndEvent): out.append("-SEQ")
            elif isinstance(ev, yaml.ScalarEvent):
                if ev.tag is not None or ev.anchor is not None:
                    raise yaml.YAMLError("tag/anchor")
                st = {None: "p", "": "p", "'": "s", '"': "d", "|": "l", ">": "f"}[ev.style]
                out.append("=VAL %d%d%s %s" % (ev.implicit[0], ev.implicit[1], st, _esc(ev.value)))
            else:
                raise yaml.YAMLError("alias")
    except yaml.YAMLError:
        err = 1
    return err, "".join(l + "\n" for l in out)


def libyaml_emit(listing: str, width=None):
    """feed a listing to libyaml's emitter"""
    evs = []
    for l in listing.split("\n"):
        if l.startswith("+STR"): evs.append(yaml.StreamStartEvent())
        elif l.startswith("-STR"): evs.append(yaml.StreamEndEvent())
        elif l.startswith("+DOC"): evs.append(yaml.DocumentStartEvent(explicit=l[4:8] == " ---"))
        elif l.startswith("-DOC"): evs.append(yaml.DocumentEndEvent(explicit=l[4:8] == " ..."))
        elif l.startswith("+MAP"): evs.append(yaml.MappingStartEvent(None, None, True, flow_style=l[4:7] == " {}"))
        elif l.startswith("-MAP"): evs.append(yaml.MappingEndEvent())
        elif l.startswith("+SEQ"): evs.append(yaml.SequenceStartEvent(None, None, True, flow_style=l[4:7] == " []"))
        elif l.startswith("-SEQ"): evs.append(yaml.SequenceEndEvent())
        elif l.startswith("=VAL "):
            p, q, s = l[5] == "1", l[6] == "1", l[7]
            raw = l[9:]
            v = []
            i = 0
            while i < len(raw):
                if raw[i] == "\\" and i + 1 < len(raw):
                    i += 1
                    v.append({"n": "\n", "r": "\r", "t": "\t", "0": "\0"}.get(raw[i], raw[i]))
                else:
                    v.append(raw[i])
                i += 1
            style = {"a": None, "p": "", "s": "'", "d": '"'}[s]
            evs.append(yaml.ScalarEvent(None, None, (p, q), "".join(v), style=style))
    return yaml.emit(evs, Dumper=yaml.CDumper, width=width).encode()
