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
shard = IslandShard(w.physics_pipeline, dist, rank, ws, torch.device("cpu"), overlap={overlap!r})
assert sorted(set(shard.owner.tolist())) == [-1, 0, 1], set(shard.owner.tolist())
assert shard.inplace == {inplace!r}, shard.inplace
for _ in range(25):
    w.physics_pipeline.step(scene.gravity, 1)
    shard.exchange()
shard.finish()
pose, vel = w.body_states()
np.save({out!r} + f"_{{rank}}.npy", np.concatenate([pose, vel], axis=1))
dist.destroy_process_group()
'''


import pytest


@pytest.mark.parametrize("shape,inplace,overlap", [((2, 2, 6), True, False), ((1, 3, 6), False, False), ((2, 2, 6), True, True), ((1, 3, 6), False, True)],
                         ids=["equal_contiguous_shards_in_place", "unequal_shards_packed", "in_place_double_buffered_async", "packed_async"])
def test_sharded_run_matches_single_process(tmp_path, shape, inplace, overlap):
    import emul_lib
    from rapier_b200 import scenes
    from rapier_b200.world import PhysicsWorld
    emul_lib.lib()
    out = str(tmp_path / "state")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out, shape=shape, inplace=inplace, overlap=overlap))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", "29533", str(script)], env=env, timeout=600)
    scene = scenes.pyramids(*shape)
    ref = PhysicsWorld(scene, _lib=emul_lib.lib())
    ref.step(25)
    pose, vel = ref.body_states()
    expect = np.concatenate([pose, vel], axis=1)
    for r in range(2):
        got = np.load(out + f"_{r}.npy")
        assert (got.view(np.uint32) == expect.view(np.uint32)).all(), f"rank {r} differs from the single-process run"


def test_partition_is_balanced_and_whole_components():
    from rapier_b200.sharding import partition_components
    comp = np.array([-1] + [1] * 55 + [56] * 55 + [111] * 55 + [166] * 55)
    owner = partition_components(comp, 2)
    assert owner[0] == -1
    for root in (1, 56, 111, 166):
        assert len(set(owner[comp == root].tolist())) == 1
    assert (owner == 0).sum() == (owner == 1).sum() == 110
