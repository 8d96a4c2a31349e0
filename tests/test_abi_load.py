"""The C-ABI library loads and exports every symbol include/bgreat_b200.h declares (no GPU needed)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bgreat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bgr_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from bgreat_b200 import _abi as A
    L = A.lib()
    declared = _declared_symbols()
    assert declared, "no declarations found"
    for s in declared:
        assert hasattr(L, s), s
    assert sorted(A.SYMBOLS) == declared
    assert L.bgr_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device context creation must fail loudly, never degrade to a host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bgreat_b200
    with pytest.raises(bgreat_b200.BgrError):
        bgreat_b200.Aligner(31)


def test_product_does_not_reference_oracle_or_emulator():
    pkg = os.path.join(ROOT, "bgreat_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle/" not in src and "liboracle" not in src and "bgr_emu" not in src, f
                assert "import oracle" not in src and "from oracle" not in src, f


def test_host_packer_through_the_abi_needs_no_gpu():
    """bgr_pack_reads is host code: the product library's own entry point, checked against plain Python
    (the device side of the packed upload is covered by tests/test_gpu_packed_upload.py)."""
    import ctypes as C
    import random

    import numpy as np
    from bgreat_b200 import _abi as A
    L = A.lib()
    rng = random.Random(7)
    reads = []
    for i in range(5000):
        ln = rng.choice([0, 1, 31, 32, 33, 100, 150, 151])
        r = "".join(rng.choice("ACGT") for _ in range(ln))
        if ln and rng.random() < 0.01:
            p = rng.randrange(ln)
            r = r[:p] + rng.choice("Nn*") + r[p + 1:]
        reads.append(r)
    text = "".join(reads).encode()
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    buf = np.frombuffer(text + b"\0", dtype=np.uint8)
    for threads in (1, 3, 8):
        words = np.full((len(text) + 31) // 32 + 1, 0xDEADBEEFDEADBEEF, dtype=np.uint64)
        bad = np.full((len(reads) + 31) // 32 + 1, 0xFFFFFFFF, dtype=np.uint32)
        nb = C.c_uint32()
        assert L.bgr_pack_reads(buf.ctypes.data, off.ctypes.data, len(reads), threads, words.ctypes.data, bad.ctypes.data,
                                C.byref(nb)) == 0
        flagged = {i for i in range(len(reads)) if (int(bad[i >> 5]) >> (i & 31)) & 1}
        assert flagged == {i for i, r in enumerate(reads) if any(c not in "ACGT" for c in r)}
        assert nb.value == len(flagged) > 5
        dirty = set()
        for i in flagged:
            dirty.update(range(int(off[i]) // 32, (int(off[i + 1]) + 31) // 32))
        code = {"A": 0, "C": 1, "G": 2, "T": 3}
        s = text.decode()
        for w in range((len(s) + 31) // 32):
            if w in dirty:
                continue
            v = 0
            for j, c in enumerate(s[32 * w:32 * w + 32]):
                v |= code[c] << (62 - 2 * j)
            assert int(words[w]) == v, w
    # argument checks
    assert L.bgr_pack_reads(buf.ctypes.data, None, 1, 1, words.ctypes.data, bad.ctypes.data, None) == 2  # BGR_ERR_ARG


def _prototypes():
    """-> {name: (return type, [parameter types])} parsed from include/bgreat_b200.h"""
    text = open(os.path.join(ROOT, "include", "bgreat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9 \*]*?)\b(bgr_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        plist = []
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                is_ptr = "*" in p or "[" in p
                base = re.sub(r"\b(const|struct)\b", "", p)
                base = re.sub(r"\[.*\]", "", base).replace("*", " ").split()
                plist.append((base[0], is_ptr))
        out[name] = (ret, plist)
    return out


def test_ctypes_signatures_match_the_header():
    """bgreat_b200/_abi.py declares every signature by hand: number of parameters, pointer-ness and integer width
    of each, and the return type are checked against the prototypes of include/bgreat_b200.h; so are the two structs."""
    import ctypes as C
    from bgreat_b200 import _abi as A
    L = A.lib()
    width = {"int": C.c_int, "uint32_t": C.c_uint32, "uint64_t": C.c_uint64, "size_t": C.c_size_t, "int32_t": C.c_int32}
    protos = _prototypes()
    assert sorted(protos) == sorted(A.SYMBOLS)
    for name, (ret, params) in protos.items():
        fn = getattr(L, name)
        if params:
            assert fn.argtypes is not None and len(fn.argtypes) == len(params), name
            for i, ((base, is_ptr), ct) in enumerate(zip(params, fn.argtypes)):
                ct_is_ptr = ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "contents") or issubclass(ct, C._Pointer)
                assert ct_is_ptr == is_ptr, (name, i, base)
                if not is_ptr:
                    assert C.sizeof(ct) == C.sizeof(width[base]), (name, i, base)
        if ret == "void":
            assert fn.restype is None, name
        elif "char" in ret:
            assert fn.restype is C.c_char_p, name
        elif ret == "uint64_t":
            assert fn.restype is C.c_uint64, name
        else:
            assert ret == "int" and fn.restype in (C.c_int, None.__class__) or fn.restype is C.c_int, name
    text = open(os.path.join(ROOT, "include", "bgreat_b200.h")).read()
    for sname, ctype in (("bgr_params", A.Params),):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (sname, sname), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = re.findall(r"(uint32_t|uint64_t|int32_t)\s+(\w+)\s*;", body)
        assert [f for _, f in fields] == [f for f, _ in ctype._fields_]
        assert C.sizeof(ctype) == sum(C.sizeof(width[t]) for t, _ in fields)
    body = re.search(r"typedef struct bgr_result \{(.*?)\} bgr_result;", text, flags=re.S).group(1)
    assert re.findall(r"uint32_t\s+(\w+)\s*;", re.sub(r"/\*.*?\*/", "", body, flags=re.S)) == ["flags_len", "path_off"]
