"""Shared by the CPU (emulated kernels) and GPU suites: incremental changes of the sets against the oracle --
the sphere-every-5-steps protocol of examples3d/b3d_large_world.rs:55-77, and body removal."""
import numpy as np

import oracle_lib
from parity_util import compare_worlds, is_exact
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld


class OracleSets:
    """The oracle behind PhysicsWorld's insert / remove / step surface."""

    def __init__(self, scene, threads=1):
        self.o = oracle_lib.OracleWorld(scene, threads=threads)
        self.nb = len(scene.bodies)
        self.nc = len(scene.colliders)

    def insert(self, body_builder, collider_builder):
        self.o.insert([body_builder.build_desc()], [collider_builder.build_desc(self.nb)])
        self.nb += 1
        self.nc += 1
        return self.nb - 1

    def remove(self, handle):
        self.o.remove_bodies([handle])

    def step(self, n=1):
        self.o.step(n)

    def body_states(self):
        return self.o.body_states()

    def debug_read(self, table, dtype):
        return self.o.debug_read(table, dtype)

    def sleeping(self):
        return self.o.sleeping()

    def set_body_states(self, handles, pose7=None, vel6=None):
        self.o.set_body_states(handles, pose7, vel6)


def large_world_protocol_case(lib=None, grid=20, spheres=12, steps=100, every=10, threads=1):
    floor = scenes.large_world_floor(grid)
    w = PhysicsWorld(scenes.large_world_floor(grid), _lib=lib)
    w.reserve(spheres, grid * grid + spheres)
    o = OracleSets(floor, threads)
    state = {"w": 0, "o": 0}

    def check(world_key):
        def f(step):
            state[world_key] = step
        return f
    # run both in lock step so that a mismatch is reported at the step it appears
    dropped = 0
    for step in range(steps):
        if dropped < spheres and step > 0 and step % 5 == 0:
            bb, cb = scenes.large_world_sphere(dropped, grid, spheres)
            w.insert(bb, cb)
            bb, cb = scenes.large_world_sphere(dropped, grid, spheres)
            o.insert(bb, cb)
            dropped += 1
        w.step()
        o.step()
        if step % every == every - 1:
            d = compare_worlds(w, o)
            assert is_exact(d), (step, d)
    assert dropped == min(spheres, (steps - 1) // 5)
    return w, o


def removal_case(lib=None):
    """A body is pulled out of a pyramid mid-run (RigidBodySet::remove): its contacts end, the rest keeps its warm
    start and collapses the same way on both sides; later a new body is dropped onto the pile."""
    from rapier_b200.sets import ColliderBuilder, RigidBodyBuilder
    scene = scenes.pyramids(1, 2, 6)
    w = PhysicsWorld(scenes.pyramids(1, 2, 6), _lib=lib)
    w.reserve(len(scene.bodies) + 4, len(scene.colliders) + 4)
    o = OracleSets(scene)
    for step in range(90):
        if step == 20:
            w.remove(3); o.remove(3)           # a base cube of the first pyramid
        if step == 30:
            w.remove(30); o.remove(30)
        if step == 40:
            for world in (w, o):
                world.insert(RigidBodyBuilder.dynamic().translation((0.3, 9.0, 0.1)).can_sleep(False), ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(100.0))
        w.step(); o.step()
        if step % 10 == 9 or step in (20, 21, 40, 41):
            d = compare_worlds(w, o)
            assert is_exact(d), (step, d)
    pose, _ = w.body_states()
    assert np.isfinite(pose).all()
