// Reference-style usage of the C++ host mirror: examples3d/b3d_many_pyramids.rs create_small_pyramid +
// `world.step()`, reading like the reference's own harness (examples3d/harness_capsules3.rs:55-62).
#include <cstdio>
#include <cstdlib>
#include "../../include/rapier_b200.hpp"

using namespace rapier_b200;

static void create_small_pyramid(PhysicsWorld& world, int base_count, float extent, float center_x, float base_z) {
    for (int i = 0; i < base_count; ++i) {
        float y = (2.0f * i + 1.0f) * extent;
        for (int j = i; j < base_count; ++j) {
            float x = (i + 1.0f) * extent + 2.0f * (j - i) * extent + center_x - 0.5f;
            world.insert(RigidBodyBuilder::dynamic().translation({x, y, base_z}).can_sleep(false),
                         ColliderBuilder::cuboid(extent, extent, extent).density(100.0f));
        }
    }
}

int main(int argc, char** argv) {
    int steps = argc > 1 ? atoi(argv[1]) : 100;
    try {
        PhysicsWorld world;
        world.gravity = {0.0f, -10.0f, 0.0f};
        world.insert(RigidBodyBuilder::fixed().translation({0.0f, -1.0f, 0.0f}), ColliderBuilder::cuboid(30.0f, 1.0f, 30.0f));
        create_small_pyramid(world, 10, 0.5f, 0.0f, 0.0f);
        create_small_pyramid(world, 10, 0.5f, 12.0f, 0.0f);
        // a bar on a limited revolute joint (revolute_joint.rs + generic_joint.rs limits), far from the pyramids
        RigidBodyHandle anchor = world.bodies.insert(RigidBodyBuilder::fixed().translation({-40.0f, 20.0f, 0.0f}));
        RigidBodyHandle bar = world.insert(RigidBodyBuilder::dynamic().translation({-39.0f, 20.0f, 0.0f}).can_sleep(false).dominance_group(3),
                                           ColliderBuilder::cuboid(1.0f, 0.1f, 0.1f));
        world.impulse_joints.insert(anchor, bar, RevoluteJointBuilder({0.0f, 0.0f, 1.0f}).local_anchor2({-1.0f, 0.0f, 0.0f}).limits(3, -0.5f, 0.25f));
        world.step(steps);
        const RbBodyDesc& bd = world.bodies.bodies[bar.index];
        const float bar_angle = 2.0f * std::atan2(bd.rotation[2], bd.rotation[3]);   // rotation about Z
        if (steps >= 100 && !(bar_angle > -0.55f && bar_angle < -0.45f)) { fprintf(stderr, "bar angle %f outside its limit\n", bar_angle); return 4; }
        RbCounters c = world.physics_pipeline.counters();
        float top_y = world.bodies.bodies[55].translation[1];   // apex cube of the first pyramid
        printf("steps=%d bodies=%d pairs=%d manifolds=%d apex_y=%.4f\n", steps, c.num_bodies, c.num_pairs, c.num_active_manifolds, top_y);
        return (c.num_active_manifolds == 290 && top_y > 9.3f && top_y < 9.6f) ? 0 : 2;
    } catch (const Error& e) {
        fprintf(stderr, "%s\n", e.what());
        return e.code == RB_ERR_NO_DEVICE ? 3 : 1;
    }
}
