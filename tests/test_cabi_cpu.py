"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/ttb200.h declares, validates arguments, and refuses to compute
without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "ttb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ttb2_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from torchtree_b200 import _lib, build

    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.ttb2_version() == 200


def test_torch_extension_builds_loads_and_links_the_cabi():
    """The autograd Functions are a torch C++ extension over the C ABI (north_star);
    no GPU here, so only loading, linkage and argument errors can be checked."""
    import subprocess

    from torchtree_b200 import _lib, build

    ext = build.build_torch_extension()
    assert os.path.exists(ext)
    from torchtree_b200 import _ttb200_torch

    assert _ttb200_torch.abi_version() == _lib.load().ttb2_version()
    needed = subprocess.run(["readelf", "-d", ext], capture_output=True, text=True).stdout
    assert "libttb200.so" in needed and "$ORIGIN/lib" in needed
    x = torch.zeros(1, 4, dtype=torch.float64)
    # a closed (here: never created) engine raises instead of dereferencing its handle
    ref = _ttb200_torch.EngineRef(0)
    assert ref.closed
    with pytest.raises(RuntimeError, match="has been closed"):
        _ttb200_torch.log_likelihood_eigen(ref, x, x, x, x, x)
    with pytest.raises(TypeError):
        _ttb200_torch.log_likelihood_eigen(0, x, x, x, x, x)   # raw handles are not accepted
    with pytest.raises(RuntimeError, match="null node-height plan"):
        _ttb200_torch.node_heights(0, 0, x)


def test_function_module_has_no_python_autograd_function():
    """function.py / height_transform.py must route through the extension, not keep a
    Python torch.autograd.Function beside it."""
    for name in ("function.py", "height_transform.py"):
        text = open(os.path.join(REPO, "torchtree_b200", name)).read()
        assert "autograd.Function)" not in text, name


def test_library_is_sm100a_only():
    import subprocess

    from torchtree_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_create_validates_arguments_before_touching_the_gpu():
    from torchtree_b200 import Engine, EngineError
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(6, 10, 4, 2, seed=1)
    with pytest.raises(EngineError, match="postorder"):
        Engine(prob.tip_states, prob.weights, prob.postorder[:-1], 4, 2)
    bad_codes = np.eye(4)[:3]
    with pytest.raises(EngineError):
        Engine(prob.tip_states, prob.weights, prob.postorder, 4, 2, code_partials=bad_codes)
    table = np.concatenate([np.eye(4), np.ones((1, 4))])[::-1].copy()  # wrong row order
    with pytest.raises(EngineError, match="unit vectors"):
        Engine(prob.tip_states, prob.weights, prob.postorder, 4, 2, code_partials=table)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from torchtree_b200 import Engine, EngineError
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(6, 10, 4, 2, seed=1)
    with pytest.raises(EngineError, match="no CPU fallback"):
        Engine(prob.tip_states, prob.weights, prob.postorder, 4, 2)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "torchtree_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_codes_from_tip_partials_round_trip():
    from torchtree_b200 import codes_from_tip_partials

    S = 4
    cols = {"A": [1, 0, 0, 0], "C": [0, 1, 0, 0], "G": [0, 0, 1, 0], "T": [0, 0, 0, 1],
            "-": [1, 1, 1, 1], "R": [1, 0, 1, 0], "Y": [0, 1, 0, 1]}
    seqs = ["ACGT-RYA", "TTGCA-RR"]
    partials = [np.array([cols[c] for c in s], dtype=np.float64).T for s in seqs]
    codes, table = codes_from_tip_partials(partials, S)
    assert codes.dtype == np.uint8 and codes.shape == (2, 8)
    assert np.array_equal(table[:5], np.concatenate([np.eye(4), np.ones((1, 4))]))
    for t, s in enumerate(seqs):
        for i, c in enumerate(s):
            assert np.array_equal(table[codes[t, i]], cols[c])
    assert table.shape[0] == 7


def test_synthetic_postorder_convention():
    from torchtree_b200.synthetic import make_problem

    for topo in ("random", "caterpillar", "balanced"):
        prob = make_problem(37, 5, topology=topo, seed=3)
        post = prob.postorder
        T = 37
        assert post.shape == (T - 1, 3)
        assert list(post[:, 0]) == list(range(T, 2 * T - 1))  # post-order numbering
        seen = set(range(T))
        for node, l, r in post:
            assert l in seen and r in seen
            seen.add(node)
        assert post[-1, 0] == 2 * T - 2
        assert 2 * T - 3 in post[-1, 1:]  # node 2T-3 is a child of the root
        assert prob.branch_lengths[0, -1] == 0.0


def test_every_entry_point_is_documented_with_its_reference_counterpart():
    """include/ttb200.h is the drop-in boundary: INTEGRATION.md must name every symbol it declares,
    and the header must cite reference files (file:line) for what the entry points replace."""
    declared = _declared_symbols()
    integration = open(os.path.join(REPO, "INTEGRATION.md")).read()
    missing = [s for s in declared if s not in integration]
    assert not missing, missing
    header = open(os.path.join(REPO, "include", "ttb200.h")).read()
    cites = re.findall(r"[a-z_/]+\.py:\d+", header)
    assert len(cites) >= 15, cites


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/ttb200.h must compile as C99 and a C program must link
    against libttb200.so and get error codes (not crashes) for bad arguments."""
    import shutil
    import subprocess

    from torchtree_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc unavailable")
    src = tmp_path / "cabi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "ttb200.h"
int main(void) {
  ttb2_engine* e = NULL;
  ttb2_config cfg;
  memset(&cfg, 0, sizeof cfg);
  if (ttb2_version() != TTB2_VERSION) return 1;
  if (ttb2_create(NULL, NULL, NULL, NULL, NULL, &e) == TTB2_OK) return 2;
  if (ttb2_last_error() == NULL || strlen(ttb2_last_error()) == 0) return 3;
  if (ttb2_loglik_q(NULL, 1, NULL, NULL, 1, NULL, 1, NULL, 1, NULL, 1, NULL, TTB2_HOST) != TTB2_E_INVALID) return 4;
  if (ttb2_coalescent_constant(0, 0, 1, NULL, NULL, 1, NULL, NULL, NULL, TTB2_HOST) != TTB2_E_INVALID) return 5;
  if (ttb2_get_config(NULL, &cfg) != TTB2_E_INVALID) return 6;
  if (ttb2_eval_serial(NULL) != 0 || ttb2_launch_count(NULL) != 0) return 7;
  ttb2_destroy(NULL);
  puts("c-abi ok");
  return 0;
}
''')
    exe = tmp_path / "cabi"
    libdir = os.path.dirname(_lib.lib_path())
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic",
                        "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lttb200", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "c-abi ok" in run.stdout, (run.returncode, run.stdout, run.stderr)
