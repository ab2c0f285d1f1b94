"""C-ABI checks that need no GPU: the library loads, exports every symbol include/rapier_b200.h
declares, agrees on struct layouts, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from rapier_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rapier_b200.h")).read()
    return sorted(set(re.findall(r"\b(rb_[a-z_0-9]+)\s*\(", text)))


def test_exports_every_declared_symbol(built):
    from rapier_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"librapier_b200.so does not export {n}"
    assert L.rb_abi_version() == 6   # (2: RbColliderDesc events fields; 3: RbJointDesc limits + motors; 4: convex polyhedra; 5: coupled_axes, additional_solver_iterations; 6: sensors, RbCollisionEvent.flags)


def test_default_parameters_match_reference_defaults(built):
    from rapier_b200 import _lib
    L = _lib.lib()
    p = A.RbIntegrationParameters()
    L.rb_integration_parameters_default(C.byref(p))
    d = A.RbIntegrationParameters.default()
    for name, _ in A.RbIntegrationParameters._fields_:
        assert getattr(p, name) == pytest.approx(getattr(d, name)), name
    # integration_parameters.rs:379-407
    assert p.num_solver_iterations == 4 and p.num_internal_pgs_iterations == 1
    assert p.normalized_prediction_distance == pytest.approx(0.02)


def test_no_cpu_fallback(built):
    """Without a CUDA device world creation must fail loudly (RB_ERR_NO_DEVICE semantics)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rapier_b200.world import PhysicsPipeline, RapierError
    with pytest.raises(RapierError) as e:
        PhysicsPipeline()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
