"""CPU tier: the C-ABI library loads and exports every symbol include/bfilt.h declares; argument validation
that needs no GPU behaves as the header says.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bayesianfiltering_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bfilt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z_0-9]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 17
    bound = {name for name, _, _ in _lib.SYMBOLS}
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in bfilt.h but not exported by libbfilt.so"
        assert name in bound, f"{name} declared in bfilt.h but not bound in _lib.SYMBOLS"
    assert lib.bf_abi_version() == 4


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "bfilt.h")).read()
    assert "torch" not in src.lower() and "at::" not in src and "#include <cuda" not in src


def _desc(**kw):
    from bayesianfiltering_b200 import zoo
    return zoo.ungm().host_model(**kw)


def test_create_rejects_bad_arguments_without_touching_the_gpu():
    lib = _lib.load()
    hm = _desc()
    h = C.c_void_p()
    cfg = _lib.PfConfig(0, 1, 0.5, 0, 0, 0, 0, 0)                 # N = 0
    assert lib.bf_pf_create(C.byref(hm.desc), C.byref(cfg), C.byref(h)) == _lib.BF_EINVAL
    assert b"num_particles" in lib.bf_last_error()
    cfg = _lib.PfConfig(100, 1, 0.5, 7, 0, 0, 0, 0)               # unknown resampler
    assert lib.bf_pf_create(C.byref(hm.desc), C.byref(cfg), C.byref(h)) == _lib.BF_EINVAL
    bad = _lib.HostModel(99, 1, 1, 1, 1, [0.0], [0], [[1]], [0], [[1]], [0], [[1]])
    cfg = _lib.PfConfig(100, 1, 0.5, 0, 0, 0, 0, 0)
    assert lib.bf_pf_create(C.byref(bad.desc), C.byref(cfg), C.byref(h)) == _lib.BF_ENOMODEL
    gcfg = _lib.GsfConfig(5000, 1, 0, 0)                          # M > 1024
    assert lib.bf_gsf_create(C.byref(hm.desc), C.byref(gcfg), C.byref(h)) == _lib.BF_EINVAL
    with pytest.raises(ValueError):
        _lib.check(_lib.BF_EINVAL)


def test_wrong_theta_length_is_rejected():
    lib = _lib.load()
    bad = _lib.HostModel(_lib.MODEL_UNGM, 1, 1, 1, 1, [0.5, 25.0], [0], [[1]], [0], [[1]], [0], [[1]])
    h = C.c_void_p()
    cfg = _lib.PfConfig(100, 1, 0.5, 0, 0, 0, 0, 0)
    assert lib.bf_pf_create(C.byref(bad.desc), C.byref(cfg), C.byref(h)) == _lib.BF_EINVAL


def test_ctypes_structs_have_the_layout_of_the_header(tmp_path):
    """include/bfilt.h is plain C: compile it with gcc, print sizeof / offsetof of every struct field, and hold the ctypes
    mirror in _lib.py (what gaussfiltax would bind) to it -- same field names, order, offsets and sizes."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = {"bf_model_desc": _lib.ModelDesc, "bf_pf_config": _lib.PfConfig, "bf_pf_io": _lib.PfIO, "bf_gsf_config": _lib.GsfConfig,
             "bf_gsf_io": _lib.GsfIO, "bf_agsf_config": _lib.AgsfConfig, "bf_agsf_io": _lib.AgsfIO}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "bfilt.h")}"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu %zu\\n", offsetof({cname}, {fname}), sizeof((({cname}*)0)->{fname}));')
    lines += [f'  printf("abi {{}} %d\\n", BF_ABI_VERSION);'.replace("{}", "version"), "  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)     # also: the header is valid C99
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = {}
    for ln in out:
        p = ln.split()
        if len(p) == 3 and p[1] == "size":
            assert C.sizeof(pairs[p[0]]) == int(p[2]), ln
        elif len(p) == 4 and p[0] in pairs:
            f = getattr(pairs[p[0]], p[1])
            assert (f.offset, f.size) == (int(p[2]), int(p[3])), ln
            seen.setdefault(p[0], []).append(p[1])
        elif len(p) == 3 and p[0] == "abi":
            assert int(p[2]) == 4
    # every field of the header is mirrored (a field the ctypes class lacks would shift nothing but lose an argument)
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bfilt.h")).read(), flags=re.S)
    for cname, cls in pairs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names += [re.sub(r"[\s\*]", "", n).split("[")[0] for n in re.sub(r"^(const\s+)?\w+\s*\*?", "", decl, count=1).split(",")]
        assert names == [n for n, _ in cls._fields_] == seen[cname], (cname, names)


def test_the_stub_in_integration_md_is_current():
    """INTEGRATION.md shows the ctypes stub a gaussfiltax maintainer would add: its structs must be those of the header
    (via the checked mirror in _lib.py) and every library symbol it names must exist."""
    import ast
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", doc, flags=re.S)
    assert len(blocks) >= 2
    ns = {}
    exec(blocks[0].replace('C.CDLL("libbfilt.so")', f"C.CDLL({_lib.LIB_PATH!r})"), ns)
    for name, mine in (("ModelDesc", _lib.ModelDesc), ("PfConfig", _lib.PfConfig), ("PfIO", _lib.PfIO)):
        theirs = ns[name]
        assert C.sizeof(theirs) == C.sizeof(mine)
        assert [(n, getattr(theirs, n).offset, getattr(theirs, n).size) for n, _ in theirs._fields_] == \
               [(n, getattr(mine, n).offset, getattr(mine, n).size) for n, _ in mine._fields_], name
    ast.parse(blocks[1])
    lib = _lib.load()
    for sym in set(re.findall(r"\b(bf_[a-z_0-9]+)\b", doc)):
        if not sym.endswith("_"):                                      # prose like `bf_gsf_*` / `bf_pf_shard_*`
            assert hasattr(lib, sym) or sym in ("bf_model_desc", "bf_pf_config", "bf_pf_io", "bf_gsf_io", "bf_gsf_config",
                                                "bf_agsf_config", "bf_agsf_io", "bf_abi_version"), sym
    assert f"BF_ABI_VERSION {lib.bf_abi_version()}" in doc
