"""N>1 host logic on CPU (gloo, world_size 2): island sharding + per-step all-gather of body states.
The worlds are the host emulation of the kernels (tests/emul); the sharded run must reproduce the
single-process run bit for bit (islands are independent, so sharding must not change any result)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import emul_lib
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld
from rapier_b200.sharding import IslandShard
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
scene = scenes.pyramids(*{shape!r})
w = PhysicsWorld(scene, _lib=emul_lib.lib()); w._flush()
shard = IslandShard(w.physics_pipeline, dist, rank, ws, torch.device("cpu"), refresh_every={refresh!r})
assert sorted(set(shard.owner.tolist())) == [-1, 0, 1], set(shard.owner.tolist())
for _ in range(25):
    w.physics_pipeline.step(scene.gravity, 1)
    shard.exchange()
assert (shard.halo_steps > 0) == {halo!r}, (shard.halo_steps, shard.export_max)
mid_pose, _ = w.body_states()     # before finish(): the rows of far foreign bodies are at most `refresh` steps old
shard.finish()
pose, vel = w.body_states()
np.save({out!r} + f"_{{rank}}.npy", np.concatenate([pose, vel], axis=1))
dist.destroy_process_group()
'''


import pytest


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("shape,halo,refresh", [((2, 2, 6), False, 8), ((1, 3, 6), True, 8), ((1, 3, 6), True, 1000)],
                         ids=["far_shards_no_per_step_collective", "neighbouring_pyramids_halo_exchange", "halo_exchange_without_refresh"])
def test_sharded_run_matches_single_process(tmp_path, shape, halo, refresh):
    import emul_lib
    from rapier_b200 import scenes
    from rapier_b200.world import PhysicsWorld
    emul_lib.lib()
    out = str(tmp_path / "state")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out, shape=shape, halo=halo, refresh=refresh))
    port = str(_free_port())
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", port, str(script)], env=env, timeout=600)
    scene = scenes.pyramids(*shape)
    ref = PhysicsWorld(scene, _lib=emul_lib.lib())
    ref.step(25)
    pose, vel = ref.body_states()
    expect = np.concatenate([pose, vel], axis=1)
    for r in range(2):
        got = np.load(out + f"_{r}.npy")
        assert (got.view(np.uint32) == expect.view(np.uint32)).all(), f"rank {r} differs from the single-process run"


def test_contact_across_shards_is_reported():
    """A body simulated here that comes to TOUCH a body simulated by another rank (here: a tracked halo body that is
    never updated) must raise RB_ERR_SHARD instead of being solved against a frozen neighbour."""
    import emul_lib
    from rapier_b200 import scenes
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    from rapier_b200.world import PhysicsWorld, RapierError
    s = scenes.Scene("two_shards", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(20.0, 0.5, 20.0))
    a = s.insert(RigidBodyBuilder.dynamic().translation((0.0, 0.5, 0.0)).linvel((8.0, 0.0, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5).friction(0.0))
    b = s.insert(RigidBodyBuilder.dynamic().translation((3.0, 0.5, 0.0)), ColliderBuilder.cuboid(0.5, 0.5, 0.5))
    w = PhysicsWorld(s, _lib=emul_lib.lib())
    w._flush()
    pipe = w.physics_pipeline
    owned = np.zeros(3, np.uint8); owned[a] = 1
    pipe.set_owned_bodies(owned)
    halo = np.zeros(3, np.uint8); halo[b] = 1
    pipe.set_halo_bodies(halo.ctypes.data)
    with pytest.raises(RapierError, match="-6"):
        for _ in range(120):
            pipe.step(s.gravity, 1)


def test_partition_is_balanced_and_whole_components():
    from rapier_b200.sharding import partition_components
    comp = np.array([-1] + [1] * 55 + [56] * 55 + [111] * 55 + [166] * 55)
    owner = partition_components(comp, 2)
    assert owner[0] == -1
    for root in (1, 56, 111, 166):
        assert len(set(owner[comp == root].tolist())) == 1
    assert (owner == 0).sum() == (owner == 1).sum() == 110
