"""CPU: the C-ABI library loads without a GPU and exports every symbol include/b200_distill.h declares; the ctypes
mirrors of the structs have the C sizes (checked with a gcc-compiled probe). No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_distill.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from dinov2_distillation_b200 import _lib
    lib = _lib.load()
    assert lib.b200_abi_version() == _lib.ABI_VERSION == 3
    declared = _declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    unbound = [s for s in declared if s not in _lib.SIGNATURES]
    assert not unbound, f"declared in the header but not bound in _lib.SIGNATURES: {unbound}"
    stale = [s for s in _lib.SIGNATURES if s not in declared]
    assert not stale, f"bound in _lib.SIGNATURES but not declared in the header: {stale}"


def test_struct_layouts_match_the_header(tmp_path):
    from dinov2_distillation_b200 import _lib
    names = {"b200_gemm_desc": _lib.GemmDesc, "b200_attn_desc": _lib.AttnDesc, "b200_vit_block": _lib.VitBlock,
             "b200_vit_config": _lib.VitConfig, "b200_projector_params": _lib.ProjectorParams,
             "b200_projector_grads": _lib.ProjectorGrads, "b200_projector_config": _lib.ProjectorConfig}
    probe = tmp_path / "probe.c"
    body = "\n".join(f'  printf("{n} %zu\\n", sizeof({n}));' for n in names)
    probe.write_text(f'#include <stdio.h>\n#include "{HEADER}"\nint main(void) {{\n{body}\n  return 0;\n}}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(probe), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for n, cls in names.items():
        assert int(sizes[n]) == C.sizeof(cls), (n, sizes[n], C.sizeof(cls))


def test_errors_are_reported_not_swallowed():
    """Argument errors come back as a negative code + message (no device needed: validation precedes any launch)."""
    from dinov2_distillation_b200 import _lib
    lib = _lib.load()
    d = _lib.GemmDesc()
    rc = lib.b200_gemm_bf16(C.byref(d), None)
    assert rc != 0
    assert b"empty problem" in lib.b200_last_error()
    with pytest.raises(_lib.B200Error):
        _lib.check(rc, "gemm")
    a = _lib.AttnDesc()
    assert lib.b200_attention_fwd(C.byref(a), None) != 0
    assert lib.b200_patch_im2col(1, 1, 1, 15, 14, 592, None) != 0
    assert b"multiple of the 14-pixel patch" in lib.b200_last_error()


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(f'#include "{HEADER}"\nint main(void) {{ return B200_ABI_VERSION == 3 ? 0 : 1; }}\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", str(src)], check=True)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under dinov2_distillation_b200/ may import or reference it."""
    pkg = os.path.join(ROOT, "dinov2_distillation_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "oracle/" in text:
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
