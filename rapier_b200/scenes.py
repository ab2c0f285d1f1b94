"""Synthetic scene builders = the reference's benchmark scene files restated as pure functions.

All coordinates are computed in f32 with the reference's expression order (the scene files use
no RNG).  Each builder returns a `Scene` (bodies, colliders, joints, gravity).
"""
import numpy as np

from .sets import (ColliderBuilder, ColliderSet, ImpulseJointSet, RevoluteJointBuilder, RigidBodyBuilder, RigidBodySet,
                   SphericalJointBuilder)

F = np.float32


class Scene:
    def __init__(self, name, gravity=(0.0, -9.81, 0.0)):
        self.name = name
        self.gravity = tuple(float(g) for g in gravity)
        self.bodies = RigidBodySet()
        self.colliders = ColliderSet()
        self.joints = ImpulseJointSet()

    def insert(self, body_builder, collider_builder):
        """PhysicsWorld::insert (src/pipeline/physics_world.rs:184-207)."""
        h = self.bodies.insert(body_builder)
        self.colliders.insert_with_parent(collider_builder, h)
        return h


def pyramids(rows=14, cols=14, base=10, extent=0.5, density=100.0, name=None):
    """examples3d/b3d_many_pyramids.rs:9-62.  (14,14,10) is the reference file (10 780 cubes);
    (8,10,20) realises BASELINE.json's "80 pyramids x 20 levels" label (16 800 cubes)."""
    s = Scene(name or f"pyramids_{rows}x{cols}x{base}", gravity=(0.0, -10.0, 0.0))
    extent = F(extent)
    ground_extent = extent * F(cols) * (F(base) + F(1.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -1.0, 0.0)),
             ColliderBuilder.cuboid(ground_extent, 1.0, ground_extent))
    base_width = F(2.0) * extent * F(base)
    base_z = -ground_extent + F(2.0) * extent
    delta_z = (F(2.0) * (ground_extent - F(2.0) * extent) / (F(rows) - F(1.0))) if rows > 1 else F(0.0)
    if rows == 1:
        base_z = F(0.0)
    for _ in range(rows):
        for j in range(cols):
            center_x = -ground_extent + F(j) * (base_width + F(2.0) * extent) + F(2.0) * extent
            for i in range(base):
                y = (F(2.0) * F(i) + F(1.0)) * extent
                for jj in range(i, base):
                    x = (F(i) + F(1.0)) * extent + F(2.0) * F(jj - i) * extent + center_x - F(0.5)
                    s.insert(RigidBodyBuilder.dynamic().translation((x, y, base_z)).can_sleep(False),
                             ColliderBuilder.cuboid(extent, extent, extent).density(density))
        base_z = base_z + delta_z
    return s


def many_pyramids():
    return pyramids(14, 14, 10, name="b3d_many_pyramids")


def many_pyramids_label():
    return pyramids(8, 10, 20, name="b3d_many_pyramids_80x20")


def single_pyramid(base=20):
    return pyramids(1, 1, base, name=f"pyramid_{base}")


def pyramid3(height=50):
    """examples3d/stress_tests/pyramid3.rs:27-56 (brick-laid 3-D pyramid; 42 925 cubes at 50)."""
    s = Scene(f"pyramid3_{height}", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -1.0, 0.0)), ColliderBuilder.cuboid(100.0, 1.0, 100.0))
    box_size, box_separation = F(2.0), F(0.5)
    half_box = F(0.5) * box_size
    h = half_box - F(0.025)
    for i in range(height):
        brick = half_box if (i & 1) else F(0.0)
        y = F(1.0) + (box_size + box_separation) * F(i)
        for j in range(i // 2, height - (i + 1) // 2):
            for k in range(i // 2, height - (i + 1) // 2):
                x = -F(height) + (box_size + F(0.25)) * F(j) + brick
                z = -F(height) + (box_size + F(0.25)) * F(k) + brick
                s.insert(RigidBodyBuilder.dynamic().translation((x, y, z)),
                         ColliderBuilder.cuboid(h, h, h).density(1000.0))
    return s


def joint_grid(n=100):
    """examples3d/b3d_joint_grid.rs:16-52."""
    s = Scene(f"joint_grid_{n}", gravity=(0.0, -10.0, 0.0))
    handles = [0] * (n * n)
    index = 0
    for k in range(n):
        for i in range(n):
            b = RigidBodyBuilder.fixed() if i == 0 else RigidBodyBuilder.dynamic().can_sleep(False)
            b = b.translation((F(k), -F(i), 0.0))
            h = s.insert(b, ColliderBuilder.ball(0.4))
            if i > 0:
                s.joints.insert(handles[index - 1], h,
                                SphericalJointBuilder().local_anchor1((0.0, -0.5, 0.0)).local_anchor2((0.0, 0.5, 0.0)))
            if k > 0:
                s.joints.insert(handles[index - n], h,
                                SphericalJointBuilder().local_anchor1((0.5, 0.0, 0.0)).local_anchor2((-0.5, 0.0, 0.0)))
            handles[index] = h
            index += 1
    return s


def keva(blocks=5):
    """examples3d/keva3.rs:5-104 (cuboid planks; 5 320 bodies at blocks=5)."""
    s = Scene(f"keva3_{blocks}", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.1, 0.0)), ColliderBuilder.cuboid(50.0, 0.1, 50.0))
    he = (F(0.02) / F(2.0) * F(10.0), F(0.1) / F(2.0) * F(10.0), F(0.4) / F(2.0) * F(10.0))
    numy_tab = [0, 9, 13, 17, 21, 41]
    block_height = F(0.0)

    def build_block(shift, numx, numy, numz):
        dims = [he, (he[2], he[1], he[0])]
        block_width = F(2.0) * he[2] * F(numx)
        bh = F(2.0) * he[1] * F(numy)
        spacing = (he[2] * F(numx) - he[0]) / (F(numz) - F(1.0))
        nx, nz = numx, numz
        for i in range(numy):
            nx, nz = nz, nx
            dim = dims[i % 2]
            y = dim[1] * F(i) * F(2.0)
            for j in range(nx):
                x = spacing * F(j) * F(2.0) if i % 2 == 0 else dim[0] * F(j) * F(2.0)
                for k in range(nz):
                    z = dim[2] * F(k) * F(2.0) if i % 2 == 0 else spacing * F(k) * F(2.0)
                    s.insert(RigidBodyBuilder.dynamic().translation(
                        (x + dim[0] + shift[0], y + dim[1] + shift[1], z + dim[2] + shift[2])),
                        ColliderBuilder.cuboid(dim[0], dim[1], dim[2]))
        dim = (he[2], he[0], he[1])
        for i in range(int(block_width / (dim[0] * F(2.0)))):
            for j in range(int(block_width / (dim[2] * F(2.0)))):
                s.insert(RigidBodyBuilder.dynamic().translation(
                    (F(i) * dim[0] * F(2.0) + dim[0] + shift[0], dim[1] + shift[1] + bh,
                     F(j) * dim[2] * F(2.0) + dim[2] + shift[2])),
                    ColliderBuilder.cuboid(dim[0], dim[1], dim[2]))

    for i in range(blocks, 0, -1):
        numx = i
        numy = numy_tab[i]
        numz = numx * 3 + 1
        block_width = F(numx) * he[2] * F(2.0)
        build_block((-block_width / F(2.0), block_height, -block_width / F(2.0)), numx, numy, numz)
        block_height = block_height + F(numy) * he[1] * F(2.0) + he[0] * F(2.0)
    return s


def box_on_ground(shape="cuboid", height=0.5):
    """crates/rapier3d/tests/total_contact_impulse.rs: unit-mass cube / ball resting on a slab."""
    s = Scene("rest_" + shape, gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(10.0, 0.5, 10.0))
    if shape == "cuboid":
        s.insert(RigidBodyBuilder.dynamic().translation((0.0, height, 0.0)),
                 ColliderBuilder.cuboid(0.5, 0.5, 0.5).density(1.0))
    else:
        r = 0.5
        vol = 4.0 / 3.0 * np.pi * r ** 3
        s.insert(RigidBodyBuilder.dynamic().translation((0.0, height, 0.0)),
                 ColliderBuilder.ball(r).density(1.0 / vol))
    return s


def box_pile(nx=6, ny=6, nz=12, jitter=True):
    """Pile + joint chain in the spirit of crates/rapier3d/tests/simd_backend_determinism.rs:61-137:
    a deterministic falling pile of cubes (tilted, so edge/vertex contacts occur) and a 4-ball chain."""
    s = Scene(f"pile_{nx}x{ny}x{nz}", gravity=(0.0, -9.81, 0.0))
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.5, 0.0)), ColliderBuilder.cuboid(30.0, 0.5, 30.0))
    rad = F(0.5)
    shift = rad * F(2.0) + F(0.1)
    for j in range(ny):
        for i in range(nx):
            for k in range(nz):
                x = F(i) * shift - F(nx) * shift * F(0.5) + (F(0.07) * F(j % 3) if jitter else F(0))
                y = F(1.0) + F(j) * (shift + F(0.3))
                z = F(k) * shift - F(nz) * shift * F(0.5) + (F(0.05) * F((j + i) % 2) if jitter else F(0))
                b = RigidBodyBuilder.dynamic().translation((x, y, z))
                if jitter:
                    b = b.rotation((0.1 * ((i + j) % 3), 0.05 * (k % 2), 0.08 * (j % 2)))
                s.insert(b, ColliderBuilder.cuboid(rad, rad, rad))
    prev = s.insert(RigidBodyBuilder.fixed().translation((F(nx) * shift, 6.0, 0.0)), ColliderBuilder.ball(0.2))
    for i in range(4):
        h = s.insert(RigidBodyBuilder.dynamic().translation((F(nx) * shift + F(i + 1) * F(0.6), 6.0, 0.0)),
                     ColliderBuilder.ball(0.2))
        s.joints.insert(prev, h, SphericalJointBuilder().local_anchor1((0.3, 0.0, 0.0)).local_anchor2((-0.3, 0.0, 0.0)))
        prev = h
    return s


REGISTRY = {
    "b3d_many_pyramids": many_pyramids,
    "b3d_many_pyramids_80x20": many_pyramids_label,
    "pyramid3": pyramid3,
    "b3d_joint_grid": joint_grid,
    "keva3": keva,
}


def offset_com_pendulums(n=128, spacing=20.0, com_offset=1.0):
    """crates/rapier3d/tests/issue_952_simd_joint_offset_com.rs:20-44: n independent pendulums, each a dynamic
    body whose collider (and so its centre of mass) is offset from the body origin, pinned at its ORIGIN to a
    fixed base by a revolute joint about Z.  (The reference uses a capsule; a cuboid of the same extent stands
    in for it -- only the offset centre of mass matters.)"""
    s = Scene(f"offset_com_pendulums_{n}")
    s.pendulums = []
    for i in range(n):
        base_pos = (float(i) * spacing, 0.0, 0.0)
        base = s.bodies.insert(RigidBodyBuilder.fixed().translation(base_pos))
        body = s.bodies.insert(RigidBodyBuilder.dynamic().translation(base_pos))
        s.colliders.insert_with_parent(ColliderBuilder.cuboid(com_offset, 0.2, 0.2).translation((com_offset, 0.0, 0.0)), body)
        s.joints.insert(base, body, RevoluteJointBuilder((0.0, 0.0, 1.0)).contacts_enabled(False))
        s.pendulums.append((body, base_pos))
    return s


def heavy_end_chain(num=17, rad=0.2):
    """crates/rapier3d/tests/substep_chain_high_mass_ratio.rs:17-55: a chain of balls on spherical joints hanging
    from a fixed ball, the last ball 10x the radius (1000:1 mass ratio).  Gravity is the default (0, -9.81, 0);
    the chain is laid out along z, so it swings."""
    s = Scene(f"heavy_end_chain_{num}")
    s.chain_joints = []
    prev = None
    shift1 = rad * 1.1
    for i in range(num):
        ball_rad = rad * 10.0 if i == num - 1 else rad
        shift2 = ball_rad + rad * 0.1
        z = 0.0 if i == 0 else (float(i) - 1.0) * 2.0 * shift1 + shift1 + shift2
        bb = (RigidBodyBuilder.fixed() if i == 0 else RigidBodyBuilder.dynamic()).translation((0.0, 0.0, z))
        h = s.insert(bb, ColliderBuilder.ball(ball_rad))
        if prev is not None:
            a1 = (0.0, 0.0, 0.0) if i == 1 else (0.0, 0.0, shift1)
            a2 = (0.0, 0.0, -shift1 * 2.0) if i == 1 else (0.0, 0.0, -shift2)
            s.joints.insert(prev, h, SphericalJointBuilder().local_anchor1(a1).local_anchor2(a2))
            s.chain_joints.append((prev, h, a1, a2))
        prev = h
    return s


def large_world(grid=1000, spheres=100, cell=10.0):
    """examples3d/b3d_large_world.rs:14-77: a grid x grid floor of parentless (fixed) cuboid colliders and `spheres`
    dynamic balls dropped over the inner 80 % of it.  The reference inserts one ball every 5 steps; here they are
    all present from the start (scene uploads replace the whole scene), staggered in height instead."""
    s = Scene(f"large_world_{grid}_{spheres}", gravity=(0.0, -10.0, 0.0))
    cell = F(cell)
    half_span = F(0.5) * cell * F(grid)
    for i in range(grid):
        x = -half_span + (F(i) + F(0.5)) * cell
        for j in range(grid):
            z = -half_span + (F(j) + F(0.5)) * cell
            s.colliders.insert(ColliderBuilder.cuboid(F(0.5) * cell, 0.25, F(0.5) * cell).translation((x, 0.0, z)))
    side = 1
    while side * side < spheres:
        side += 1
    inset = F(0.1) * F(2.0) * half_span
    usable = F(2.0) * half_span - F(2.0) * inset
    for idx in range(spheres):
        gi, gj = idx % side, idx // side
        x = -half_span + inset + (F(gi) + F(0.5)) * (usable / F(side))
        z = -half_span + inset + (F(gj) + F(0.5)) * (usable / F(side))
        s.insert(RigidBodyBuilder.dynamic().translation((x, F(1.5) + F(0.25) * F(idx % 7), z)), ColliderBuilder.ball(0.5))
    return s


def large_world_floor(grid=1000, cell=10.0, name=None):
    """The static part of examples3d/b3d_large_world.rs:27-41: grid x grid parentless cuboid colliders."""
    s = Scene(name or f"large_world_floor_{grid}", gravity=(0.0, -10.0, 0.0))
    cell = F(cell)
    half_span = F(0.5) * cell * F(grid)
    for i in range(grid):
        x = -half_span + (F(i) + F(0.5)) * cell
        for j in range(grid):
            z = -half_span + (F(j) + F(0.5)) * cell
            s.colliders.insert(ColliderBuilder.cuboid(F(0.5) * cell, 0.25, F(0.5) * cell).translation((x, 0.0, z)))
    return s


def large_world_sphere(idx, grid=1000, spheres=100, cell=10.0):
    """The idx-th sphere of b3d_large_world.rs:55-70 (dropped every 5 steps at y = 1.5 over the inner 80 % of the floor)."""
    cell = F(cell)
    half_span = F(0.5) * cell * F(grid)
    side = 1
    while side * side < spheres:
        side += 1
    gi, gj = idx % side, idx // side
    inset = F(0.1) * F(2.0) * half_span
    usable = F(2.0) * half_span - F(2.0) * inset
    x = -half_span + inset + (F(gi) + F(0.5)) * (usable / F(side))
    z = -half_span + inset + (F(gj) + F(0.5)) * (usable / F(side))
    return RigidBodyBuilder.dynamic().translation((x, 1.5, z)), ColliderBuilder.ball(0.5)


def run_large_world_protocol(world, steps, grid=1000, spheres=100, drop_interval=5, on_step=None):
    """StepLargeWorld (b3d_large_world.rs:55-77): one sphere inserted every `drop_interval` steps (from step
    `drop_interval` on) up to `spheres`, each followed by a step.  `world` needs insert(body_builder,
    collider_builder) and step()."""
    dropped = 0
    for step_count in range(steps):
        if dropped < spheres and step_count > 0 and step_count % drop_interval == 0:
            bb, cb = large_world_sphere(dropped, grid, spheres)
            world.insert(bb, cb)
            dropped += 1
        world.step()
        if on_step is not None:
            on_step(step_count)
    return dropped


REGISTRY["large_world"] = large_world
REGISTRY["falling_pile_2000"] = lambda: box_pile(10, 10, 20)
REGISTRY["pyramid3_20"] = lambda: pyramid3(20)


def convex_polyhedra(layers=25, num=5, scale=2.0, border_rad=0.1, seed=0):
    """examples3d/convex_polyhedron3.rs:15-58: `layers` x num x num round convex hulls of 10 random points in
    [0, scale)^3 dropped on a 40 x 0.1 x 40 ground slab.  The reference draws the points from rand's StdRng(0), which
    cannot be reproduced here: numpy's default_rng(seed) instead (same distribution, other points)."""
    s = Scene(f"convex_polyhedra_{layers}")
    s.insert(RigidBodyBuilder.fixed().translation((0.0, -0.1, 0.0)), ColliderBuilder.cuboid(40.0, 0.1, 40.0))
    shift = F(border_rad) * F(2.0) + F(scale)
    centerx = shift * F(num // 2)
    centery = shift / F(2.0)
    centerz = shift * F(num // 2)
    rng = np.random.default_rng(seed)
    for j in range(layers):
        for i in range(num):
            for k in range(num):
                x = F(i) * shift - centerx
                y = F(j) * shift + centery + F(3.0)
                z = F(k) * shift - centerz
                pts = rng.random((10, 3), dtype=np.float32) * F(scale)
                s.insert(RigidBodyBuilder.dynamic().translation((x, y, z)), ColliderBuilder.round_convex_hull(pts, border_rad))
    return s
