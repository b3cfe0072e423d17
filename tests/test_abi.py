"""The C-ABI library loads and exports every symbol include/sim_juncs_b200.h declares (no compute)."""
import ctypes
import os
import re

from sim_juncs_b200 import _lib


def declared_symbols(root):
    hdr = open(os.path.join(root, "include", "sim_juncs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sj_[a-z_0-9]+)\s*\(", hdr)))


def test_header_matches_binding(root):
    names = declared_symbols(root)
    assert len(names) >= 20
    assert sorted(_lib.SYMBOLS) == names


def test_library_exports_every_symbol(root):
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols(root):
        assert hasattr(L, name), name
    assert _lib.load().sj_version() >= 100


def test_struct_layouts_match_header(root, tmp_path):
    # sizes the C compiler produces for the header's structs must equal the ctypes mirrors
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "sim_juncs_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(sj_grid),sizeof(sj_pole),sizeof(sj_material),sizeof(sj_csg_node),sizeof(sj_region));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(t) for t in (_lib.SjGrid, _lib.SjPole, _lib.SjMaterial, _lib.SjCsgNode, _lib.SjRegion)]


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from sim_juncs_b200 import Sim, SjError
    with pytest.raises(SjError):
        Sim((8, 8, 8), 4.0)
