"""The C-ABI library on a machine without a GPU: it loads, exports every symbol the header declares,
its POD structs have the layout the ctypes mirror assumes, and it refuses to compute (no CPU fallback)."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "ocd_b200.h"


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ocd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import l4dc_mpc_ocd_b200 as ocd
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(ocd._native.lib, n), f"{n} is declared in include/ocd_b200.h but not exported"
    assert set(ocd._native.EXPORTS) == set(names)
    assert ocd._native.lib.ocd_abi_version() == ocd._native.ABI_VERSION
    assert ocd._native.strerror(0) == "ok" and "CUDA" in ocd._native.strerror(ocd._native.ECUDA)


def test_struct_layout_matches_header(tmp_path):
    """Compile a probe against the header and compare sizeof/offsetof with the ctypes structs."""
    import l4dc_mpc_ocd_b200 as ocd
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ocd_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ocd_params), '
                   'offsetof(ocd_params, lr), offsetof(ocd_params, lane_x), sizeof(ocd_scenario), '
                   'offsetof(ocd_scenario, init_state), offsetof(ocd_scenario, plan), '
                   'offsetof(ocd_scenario, teleport_state), offsetof(ocd_params, math_mode));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    P, S = ocd._native.ocd_params, ocd._native.ocd_scenario
    want = [C.sizeof(P), P.lr.offset, P.lane_x.offset, C.sizeof(S), S.init_state.offset, S.plan.offset,
            S.teleport_state.offset, P.math_mode.offset]
    assert got == want


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import l4dc_mpc_ocd_b200 as ocd
    assert ocd.device_count() == 0
    with pytest.raises(ocd._native.OcdCudaError):
        ocd.Engine(0)
    h = C.c_void_p()
    assert ocd._native.lib.ocd_ctx_create(0, C.byref(h)) == ocd._native.ECUDA
    # the drop-in classes fail the same way instead of computing on the host
    from l4dc_mpc_ocd_b200.interact_drive import simulation_utils
    with pytest.raises(ocd._native.OcdCudaError):
        simulation_utils.next_car_state([0., 0., 1., 0.], [0., 0.], 0.1)


def test_argument_validation_needs_no_device():
    """Bad arguments are rejected before anything touches CUDA."""
    import l4dc_mpc_ocd_b200 as ocd
    lib, N = ocd._native.lib, ocd._native
    p = ocd.PlannerParams(H=65).c_struct()
    one = (C.c_float * 64)()
    assert lib.ocd_solve_batch(C.addressof(p), one, None, 0, one, 1, None, None, one, one, one, None, 1, None) == N.EUNSUP
    p = ocd.PlannerParams().c_struct()
    assert lib.ocd_solve_batch(C.addressof(p), None, None, 0, one, 1, None, None, one, one, one, None, 1, None) == N.EINVAL
    assert lib.ocd_solve_batch(C.addressof(p), one, None, 0, one, 3, None, None, one, one, one, None, 8, None) == N.EINVAL
    assert lib.ocd_solve_batch(C.addressof(p), None, None, 0, None, 1, None, None, None, None, None, None, 0, None) == N.OK
    assert lib.ocd_num_starts(C.addressof(p)) == 3
    p6 = ocd.PlannerParams(extra_inits=True).c_struct()
    assert lib.ocd_num_starts(C.addressof(p6)) == 6
    with pytest.raises(ValueError):
        N.check(N.EINVAL, "x")
    with pytest.raises(MemoryError):
        N.check(N.ENOMEM)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package, include/ or csrc/ may reference it."""
    pkg = ROOT / "l4dc-mpc-ocd_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + [HEADER]:
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
        assert "ocd_oracle" not in text and "libocd_oracle" not in text, f
    out = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import l4dc_mpc_ocd_b200, "
                          "l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord, l4dc_mpc_ocd_b200.experiments.run_mpc_ord; "
                          "print('oracle' in sys.modules)" % str(ROOT)], capture_output=True, text=True, check=True)
    assert out.stdout.strip() == "False"


def test_integration_stub_matches_the_struct():
    """INTEGRATION.md shows the ctypes binding a maintainer of the reference would add; its ocd_params must have
    the fields of the real struct, in order (a stale copy would silently mis-align every argument)."""
    import l4dc_mpc_ocd_b200 as ocd
    text = (ROOT / "INTEGRATION.md").read_text()
    block = text[text.index("class ocd_params(C.Structure)"):]
    block = block[:block.index("assert _lib.ocd_abi_version()")]
    names = re.findall(r'\("(\w+)",\s*C\.c_', block)
    assert names == [f[0] for f in ocd._native.ocd_params._fields_]
    assert "ocd_abi_version() == %d" % ocd._native.ABI_VERSION in text
