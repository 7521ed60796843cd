"""The ctypes mirrors in warp_drive_b200/lib.py must have the layout of the structs in
include/wdb200.h: a C program compiled against the header prints sizeof / offsetof of every
struct and field, and the test compares them with ctypes (a mismatch would silently shift
every later argument of a launch)."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PAIRS = {
    "wdb_reset_desc": "ResetDesc", "wdb_tc_env": "TcEnv", "wdb_tc_policy_io": "TcPolicyIO",
    "wdb_tc_rollout": "TcRollout", "wdb_sa_rollout": "SaRollout",
    "wdb_gather_policy": "GatherPolicy", "wdb_gather": "Gather",
    "wdb_bookkeep_policy": "BookkeepPolicy", "wdb_bookkeep": "Bookkeep",
    "wdb_mlp_pair": "MlpPair", "wdb_pg_loss": "PgLoss",
}


def _c_fields(header, struct):
    """Field names of `typedef struct <struct> { ... } <struct>;` in declaration order."""
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (struct, struct), header, re.S)
    assert body, struct
    text = re.sub(r"/\*.*?\*/", "", body.group(1), flags=re.S)
    names = []
    for decl in text.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        # "const float *a, *b" / "int x[4]" / "wdb_tc_policy_io policy[4]" / "void *p"
        first, *rest = decl.split(",")
        parts = [first.split()[-1]] + [r.strip() for r in rest]
        for p in parts:
            names.append(re.sub(r"\[.*\]", "", p).lstrip("*").strip())
    return names


def test_ctypes_structs_match_the_header(tmp_path):
    from warp_drive_b200 import lib as wlib

    header = open(os.path.join(ROOT, "include", "wdb200.h"), encoding="utf8").read()
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "wdb200.h"', "int main(void) {"]
    fields = {}
    for c_name in PAIRS:
        fields[c_name] = _c_fields(header, c_name)
        lines.append(f'  printf("{c_name} size %zu\\n", sizeof({c_name}));')
        for f in fields[c_name]:
            lines.append(f'  printf("{c_name} {f} %zu\\n", offsetof({c_name}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for line in out.splitlines():
        s, f, v = line.split()
        got[(s, f)] = int(v)
    for c_name, py_name in PAIRS.items():
        cls = getattr(wlib, py_name)
        assert ctypes.sizeof(cls) == got[(c_name, "size")], (c_name, ctypes.sizeof(cls))
        py_fields = [n for n, _ in cls._fields_]
        assert py_fields == fields[c_name], (c_name, py_fields, fields[c_name])
        for f in py_fields:
            assert getattr(cls, f).offset == got[(c_name, f)], (c_name, f)


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/wdb200.h that lib.py binds has the same number of arguments
    there, and pointer / integer / float positions agree."""
    from warp_drive_b200 import lib as wlib

    header = open(os.path.join(ROOT, "include", "wdb200.h"), encoding="utf8").read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:^|\n)\s*(?:int|long long|const char \*)\s*(wdb_\w+)\s*\((.*?)\)\s*;",
                         header, re.S):
        args = " ".join(m.group(2).split())
        protos[m.group(1)] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    assert len(protos) > 40
    checked = 0
    for name, (_restype, argtypes) in wlib._SIGNATURES.items():
        assert name in protos, f"{name} is bound in lib.py but not declared in wdb200.h"
        c_args = protos[name]
        assert len(c_args) == len(argtypes), (name, len(c_args), len(argtypes))
        for c, t in zip(c_args, argtypes):
            is_ptr = "*" in c
            if is_ptr:
                assert (t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents")
                        or "LP_" in t.__name__), (name, c, t)
            elif re.match(r"(const )?(float)\b", c):
                assert t is ctypes.c_float, (name, c, t)
            elif re.match(r"(const )?(long long|unsigned long long)\b", c):
                assert t in (ctypes.c_longlong, ctypes.c_ulonglong), (name, c, t)
            elif re.match(r"(const )?(int|unsigned int|unsigned)\b", c):
                assert t in (ctypes.c_int, ctypes.c_uint), (name, c, t)
        checked += 1
    assert checked > 40
