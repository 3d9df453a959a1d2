"""The C-ABI library loads and exports exactly what include/b200engine.h declares; the product path
fails loudly (never falls back) when no GPU is present.  CPU only, no compute calls."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "b200engine.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from kubeai_b200 import _lib
    l = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/b200engine.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    assert l.b200_version().startswith(b"kubeai-b200")


def test_library_is_in_tree_and_self_contained():
    from kubeai_b200 import _lib
    assert _lib.LIB_PATH.exists() and ROOT in _lib.LIB_PATH.parents
    import subprocess
    deps = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libtorch" not in deps and "libpython" not in deps, "the engine must not depend on torch/python"
    assert "libcudart" in deps or "cudart" in deps


def test_struct_layouts_match_header():
    """sizeof/offset spot checks of the ctypes mirrors against the C header (compiled with gcc)."""
    import subprocess, tempfile
    from kubeai_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "b200engine.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu ", sizeof(b200_config), sizeof(b200_sampling), sizeof(b200_usage), sizeof(b200_stats), sizeof(b200_step_info), offsetof(b200_config, num_kv_blocks));
  printf("%zu %zu %zu\n", offsetof(b200_config, seed), offsetof(b200_stats, kernel_launches), offsetof(b200_step_info, device_us));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "t.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(Path(d) / "t.c"), "-o", str(Path(d) / "t")], check=True)
        got = list(map(int, subprocess.run([str(Path(d) / "t")], capture_output=True, text=True, check=True).stdout.split()))
    want = [C.sizeof(_lib.Config), C.sizeof(_lib.Sampling), C.sizeof(_lib.Usage), C.sizeof(_lib.Stats),
            C.sizeof(_lib.StepInfo), _lib.Config.num_kv_blocks.offset, _lib.Config.seed.offset,
            _lib.Stats.kernel_launches.offset, _lib.StepInfo.device_us.offset]
    assert got == want


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    from kubeai_b200 import B200Error, _lib
    from kubeai_b200.engine import Engine, mini_config
    with pytest.raises(B200Error) as ei:
        Engine(mini_config())
    assert ei.value.code == -5 and "no CPU fallback" in str(ei.value)
    l = _lib.lib()
    assert l.b200_op_gemm(None, None, None, 128, 16, 64, None) == -5


def test_product_package_never_imports_the_oracle():
    for p in (ROOT / "kubeai_b200").rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), flags=re.M), p
    for p in (ROOT / "kubeai_b200" / "csrc").glob("*"):
        assert "oracle/" not in p.read_text(errors="ignore"), p


def test_go_shim_calls_only_declared_and_exported_functions():
    """shim/engine.go (the cgo binding a KubeAI maintainer adds; no Go toolchain here) must bind real entry points."""
    go = (ROOT / "shim" / "engine.go").read_text()
    used = set(re.findall(r"C\.(b200_[a-z0-9_]+)\(", go))
    assert {"b200_engine_create", "b200_submit", "b200_poll", "b200_wait", "b200_abort", "b200_server_create",
            "b200_server_handle", "b200_engine_is_failed"} <= used
    header = (ROOT / "include" / "b200engine.h").read_text()
    from kubeai_b200 import _lib
    l = _lib.lib()
    for fn in used:
        assert re.search(r"\b%s\s*\(" % fn, header), f"{fn} is not declared in include/b200engine.h"
        assert hasattr(l, fn), f"{fn} is not exported by libb200engine.so"
    for imp in ("errors", "io", "net/http", "runtime/cgo", "unsafe"):
        assert f'"{imp}"' in go, f"import {imp} missing"
    # struct fields the shim sets exist in the header
    for field in ("max_tokens", "ignore_eos", "mean_load_pct", "prefix_char_length", "default_max_tokens"):
        assert field in header
