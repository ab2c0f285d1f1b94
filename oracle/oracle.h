// TEST INFRASTRUCTURE ONLY -- CPU oracle for the rapier step hot path (3-D, f32).
//
// PARITY UNPINNED: the reference (dimforge/rapier 0.35.2, Rust) cannot be built or run in this
// environment (no cargo/rustc), its contact-manifold geometry lives in the
// un-vendored crate parry3d 0.30.2, and its only golden vectors are whole-simulation bit hashes
// of its own binary (crates/rapier3d/tests/simd_backend_determinism.rs:140-150).  This oracle is a
// scalar restatement of the reference's algorithm, function by function with file:line citations,
// and is pinned only BEHAVIOURALLY against the reference's physical known-answer tests
// (tests/test_oracle_kat.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  The product (rapier_b200/) never links or calls it.
#pragma once
#include <stdint.h>
#include "../include/rapier_b200.h"   // POD descriptors only (RbBodyDesc, RbColliderDesc, ...)

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcWorld OrcWorld;

OrcWorld* orc_world_create(const RbIntegrationParameters* params);
void orc_world_destroy(OrcWorld* w);
int orc_world_set_params(OrcWorld* w, const RbIntegrationParameters* params);
int orc_world_set_scene(OrcWorld* w, int32_t nb, const RbBodyDesc* bodies, int32_t nc,
                        const RbColliderDesc* colliders, int32_t nj, const RbJointDesc* joints);
// Mirrors of rb_world_get_sleeping / rb_world_wake_up.
int orc_world_get_sleeping(OrcWorld* w, uint8_t* sleeping);
int orc_world_get_quarantine(OrcWorld* w, int32_t* bodies, int32_t cap);
int orc_world_wake_up(OrcWorld* w, int32_t n, const int32_t* indices);
// Mirrors of rb_world_insert / rb_world_remove_bodies (appended bodies and colliders; tombstoned removals).
int32_t orc_world_add_hull(OrcWorld* w, int32_t num_vertices, const float* vertices3, int32_t num_faces,
                           const int32_t* face_sizes, const int32_t* face_indices);   /* rb_world_add_hull */
int orc_world_insert(OrcWorld* w, int32_t nb, const RbBodyDesc* bodies, int32_t nc, const RbColliderDesc* colliders);
int orc_world_remove_bodies(OrcWorld* w, int32_t n, const int32_t* indices);
int orc_world_insert_joints(OrcWorld* w, int32_t n, const RbJointDesc* joints);      /* rb_world_insert_joints */
int orc_world_remove_joints(OrcWorld* w, int32_t n, const int32_t* indices);        /* rb_world_remove_joints */
int orc_world_update_joints(OrcWorld* w, int32_t n, const int32_t* indices, const RbJointDesc* joints, int32_t wake_up);   /* rb_world_update_joints */
int orc_world_set_body_states(OrcWorld* w, int32_t n, const int32_t* indices, const float* pose7,
                              const float* vel6);
int orc_world_step(OrcWorld* w, const float gravity[3], int32_t nsteps);
int orc_world_set_body_forces(OrcWorld* w, int32_t n, const int32_t* indices, const float* force3, const float* torque3);
int orc_world_set_next_kinematic_positions(OrcWorld* w, int32_t n, const int32_t* indices, const float* pose7);
int orc_world_drain_collision_events(OrcWorld* w, int32_t cap, RbCollisionEvent* out);
int orc_world_drain_contact_force_events(OrcWorld* w, int32_t cap, RbContactForceEvent* out);
int orc_world_get_body_states(OrcWorld* w, float* pose7, float* vel6);
int orc_world_num_bodies(OrcWorld* w);
int orc_world_get_counters(OrcWorld* w, RbCounters* out);
int orc_world_get_contact_pairs(OrcWorld* w, int32_t cap, int32_t* pair_colliders, int32_t* num_contacts,
                                int32_t* color, float* normal, float* impulses);
// Name-addressed table dump mirroring rb_world_debug_read (same table names and layouts).
int64_t orc_world_debug_read(OrcWorld* w, const char* table, void* dst, int64_t cap_bytes);
// OpenMP threads used for the colour-parallel sweeps (1 = scalar port).
void orc_set_threads(int n);
int orc_get_threads(void);

// Single-function entry points used by unit-level parity tests.
// Cuboid-cuboid / ball manifold for pose12 (pose of shape 2 in shape 1's frame).
// out_points: up to 8 * 9 floats {local_p1(3), local_p2(3), dist, fid1, fid2 (as float bits)}.
int orc_contact_manifold(int shape1, const float he1[3], int shape2, const float he2[3],
                         const float pos12_t[3], const float pos12_q[4], float prediction,
                         float* out_points, float out_n1[3], float out_n2[3]);

// One function of the path on literal inputs (see oracle_capi.cpp).
int orc_kat(const char* name, const float* in, int32_t n_in, float* out, int32_t n_out);

#ifdef __cplusplus
}
#endif
