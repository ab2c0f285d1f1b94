"""Golden fixtures (tests/golden/*.npz, generated from the oracle by tests/golden/make_golden.py):
the oracle must still reproduce them bit for bit (CPU), and so must the CUDA path through the C ABI (GPU)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same_bits(a, b):
    return a.shape == b.shape and bool((a.view(np.uint32) == b.view(np.uint32)).all())


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden(name):
    import oracle_lib
    scene, steps = make_golden.build(name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert int(g["steps"]) == steps
    o = oracle_lib.OracleWorld(scene)
    o.step(steps)
    pose, vel = o.body_states()
    assert _same_bits(pose.astype(np.float32), g["pose"]) and _same_bits(vel.astype(np.float32), g["vel"])


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_emulated_kernels_reproduce_golden(name):
    """The kernels' phase functions compiled for the host (tests/emul), i.e. the kernel LOGIC without a GPU."""
    import emul_lib
    from rapier_b200.world import PhysicsWorld
    scene, steps = make_golden.build(name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    w = PhysicsWorld(scene, _lib=emul_lib.lib())
    w.step(steps)
    pose, vel = w.body_states()
    assert _same_bits(pose.astype(np.float32), g["pose"]) and _same_bits(vel.astype(np.float32), g["vel"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_cuda_reproduces_golden(built, name):
    from rapier_b200.world import PhysicsWorld
    scene, steps = make_golden.build(name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    w = PhysicsWorld(scene)
    w.step(steps)
    pose, vel = w.body_states()
    assert _same_bits(pose.astype(np.float32), g["pose"]) and _same_bits(vel.astype(np.float32), g["vel"])
