// TEST INFRASTRUCTURE ONLY -- internal data model of the CPU oracle (see oracle.h header note).
#pragma once
#include <vector>
#include <cstdint>
#include "omath.h"
#include "oracle.h"

namespace orc {

constexpr int MAX_MANIFOLD_POINTS = 4;     // src/lib.rs:291
constexpr int MAX_RAW_POINTS = 8;          // two convex quads clip to <= 8 vertices
constexpr uint8_t COLOR_UNCOLORED = 255;   // contact_pair.rs:152
constexpr uint8_t COLOR_OVERFLOW = 128;    // contact_pair.rs:155
constexpr int DYNAMIC_COLOR_COUNT = 120;   // contact_pair.rs:159
constexpr int NUM_COLORS = 129;
constexpr uint32_t NO_BODY = 0xffffffffu;  // u32::MAX solver id = world-attached

struct Mask128 {
    uint64_t lo = 0, hi = 0;
    bool test(int c) const { return c < 64 ? ((lo >> c) & 1) : ((hi >> (c - 64)) & 1); }
    void set(int c) { if (c < 64) lo |= (1ull << c); else hi |= (1ull << (c - 64)); }
    void clear(int c) { if (c < 64) lo &= ~(1ull << c); else hi &= ~(1ull << (c - 64)); }
};

struct Params {
    RbIntegrationParameters p;
    float prediction_distance() const { return p.normalized_prediction_distance * p.length_unit; }
    float max_corrective_velocity() const {
        return p.normalized_max_corrective_velocity * p.length_unit;
    }
    float max_linear_velocity() const { return p.normalized_max_linear_velocity * p.length_unit; }
    float contact_recycle_distance() const {
        return p.normalized_contact_recycle_distance * p.length_unit;
    }
};

struct Body {
    int type;
    uint32_t flags;
    Pose pos;        // RigidBodyPosition::position
    Pose next_pos;   // RigidBodyPosition::next_position
    Pose kin_target; // ... as set by set_next_kinematic_position (position-based kinematic bodies)
    V3 linvel, angvel;
    float lin_damping, ang_damping, gravity_scale;
    float additional_mass;   // RigidBodyAdditionalMassProps::Mass (0 = none)
    V3 user_force, user_torque;
    // local mass properties (parry MassProperties)
    V3 local_com;
    float inv_mass;
    V3 principal_inertia, inv_principal_inertia;
    Q4 principal_frame;
    // world-space (RigidBodyMassProps)
    V3 world_com;
    V3 eff_inv_mass;
    Sdp3 eff_world_inv_inertia;
    // per-step forces
    V3 force, torque;
    // RigidBodyActivation (rigid_body_components.rs:1296-1326) + mprops.max_extent (:491-515)
    bool sleeping = false;
    float sleep_time = 0.0f;          // time_since_can_sleep
    Pose sleep_prev_pose{Q4{0.f, 0.f, 0.f, 1.f}, V3{0.f, 0.f, 0.f}};
    float max_extent = 0.0f;
    float ccd_thickness = 3.4028235e38f;   // RigidBodyCcd::ccd_thickness
    // Kinematic bodies are solver bodies like dynamic ones (zero effective inverse mass; solver_body.rs:112-120) and are
    // coloured like them (narrow_phase/mod.rs:105-106).  DEVIATION shared with the kernels: they are island members too.
    bool is_dynamic() const { return type == RB_BODY_DYNAMIC || type == RB_BODY_KINEMATIC_POSITION_BASED || type == RB_BODY_KINEMATIC_VELOCITY_BASED; }
    bool is_strict_dynamic() const { return type == RB_BODY_DYNAMIC; }
    // RigidBodyDominance::effective_group (rigid_body_components.rs:1267-1275)
    int effective_dominance() const { return is_dynamic() ? RB_BODY_DOMINANCE_OF(flags) : 128; }
    bool is_awake() const { return is_dynamic() && !sleeping; }   // member of the active set
};

// Convex polyhedra (parry ConvexPolyhedron as used by ColliderBuilder::convex_hull / convex_mesh); oracle_hull.cpp
constexpr int HULL_MAX_VERTS = 32, HULL_MAX_FACES = 32, HULL_MAX_FACE_VERTS = 8, HULL_MAX_EDGES = 64;
struct HullEdge { int v0, v1, f0, f1; };   // f0: the face on which the edge runs v0 -> v1
struct Hull {
    std::vector<V3> verts;
    std::vector<int> face_start, face_count, loops;   // counter-clockwise loops seen from outside
    std::vector<V3> normals;
    std::vector<float> offsets;
    std::vector<HullEdge> edges;
    float volume = 0.0f;                               // unit-density mass properties about the centre of mass
    V3 com{0.f, 0.f, 0.f}, principal_inertia{0.f, 0.f, 0.f};
    Q4 principal_frame{0.f, 0.f, 0.f, 1.f};
    V3 aabb{0.f, 0.f, 0.f};                            // max |coordinate| per axis
    float radius = 0.0f;                               // max |vertex|
};
bool hull_from_mesh(int nv, const float* verts, int nf, const int32_t* face_sizes, const int32_t* face_indices, Hull& h);
void hull_unit_cube(Hull& h);

struct Aabb {
    V3 mins, maxs;
};

struct Collider {
    int shape;
    V3 he;          // cuboid half extents, or (r,0,0)
    int parent;     // -1 = none
    Pose pos_wrt_parent;
    float density, friction, restitution;
    int friction_rule, restitution_rule;
    float contact_skin;
    uint32_t memberships, filter;
    uint32_t active_events;         // ActiveEvents
    int sensor;                     // Collider::is_sensor: intersection events only (narrow_phase/intersections.rs)
    float force_event_threshold;    // contact_force_event_threshold
    Pose pos;       // world pose
    Aabb aabb;      // compute_broad_phase_aabb (tight + skin + prediction/2)
    Aabb fat;       // broad-phase leaf AABB (change-detection skin)
    bool fat_valid;
};

struct Point {  // parry TrackedContact + rapier ContactData (contact_pair.rs:54-87)
    V3 local_p1, local_p2;
    float dist;
    uint32_t fid1, fid2;
    float impulse, warmstart_impulse, warmstart_twist;
    V3 warmstart_tangent_world;
    V3 dp1, dp2;
};

struct SolverContact {  // contact_pair.rs:617-653 (anchors already localised)
    V3 anchor1, anchor2;
    int cid;
};

struct Pair {
    int c1, c2;
    int b1, b2;            // parent bodies (-1 none)
    bool has_recycle;      // ContactRecycleState (contact_pair.rs:256-282)
    Pose r_pos12;
    Q4 r_rot1, r_rot2;
    float r_max_extent, r_max_drift;
    int npts;
    V3 local_n1, local_n2;
    Point pts[MAX_MANIFOLD_POINTS];
    V3 normal;
    float friction, restitution;
    int nsc;
    SolverContact sc[MAX_MANIFOLD_POINTS];
    uint8_t color;
    uint32_t color_bodies[2];
    bool force_event_emitted;   // PairEventStatus::INITIAL_FORCE_THRESHOLD_EVENT_EMITTED
    bool intersecting;          // IntersectionPair::intersecting of a pair with a sensor
};

// ContactManifoldData::relative_dominance (pair_update.rs:381-382): > 0 = body 1 dominates (world-attached in this contact).
inline int relative_dominance(const Body* rb1, const Body* rb2) {
    return (rb1 ? rb1->effective_dominance() : 128) - (rb2 ? rb2->effective_dominance() : 128);
}

struct RawPoint {
    V3 local_p1, local_p2;
    float dist;
    uint32_t fid1, fid2;
};
struct RawManifold {
    int n;
    RawPoint pts[MAX_RAW_POINTS];
    V3 local_n1, local_n2;
};

// Geometry (parry restatement) -- oracle_geom.cpp
void contact_manifold(int shape1, V3 he1, int shape2, V3 he2, const Pose& pos12, float prediction,
                      RawManifold& out);
Aabb shape_aabb(int shape, V3 he, const Pose& pos);
// pairs with a convex polyhedron, and its AABB -- oracle_poly.cpp
void contact_manifold_convex(const std::vector<Hull>& hulls, int sh1, V3 he1, int sh2, V3 he2, const Pose& p12, float prediction,
                             RawManifold& out);
Aabb convex_aabb(const std::vector<Hull>& hulls, V3 he, const Pose& pos);

struct Joint {
    int body1, body2;
    Pose local_frame1, local_frame2;   // as given (body space)
    Pose sframe1, sframe2;             // solver-body space (generic_joint.rs:624-636)
    uint32_t locked_axes;
    int contacts_enabled;
    float natural_frequency, damping_ratio;
    uint32_t sid1, sid2;               // solver ids (NO_BODY = world attached)
    int color;
    float impulses[6];
    // limits and motors of the free axes (generic_joint.rs:142-232, :268-300)
    uint32_t limit_axes, motor_axes;
    uint32_t coupled_axes;             // JointAxesMask of the coupled axes (generic_joint.rs: spring / rope joints couple LIN_AXES)
    float limits[6][2];
    RbJointMotor motors[6];
    float ang_limit_center[3][2], ang_limit_half_range[3];   // AngularLimitParams (joint_constraint_helper.rs:34-73)
    float limit_impulses[6], motor_impulses[6];
    bool removed;                      // ImpulseJointSet::remove: the slot stays, the joint is neither solved nor an island edge
};

struct JointRow {  // JointConstraint<Real,1> (joint_velocity_constraint.rs:68-93)
    V3 lin_jac, ang_jac1, ang_jac2, ii_ang_jac1, ii_ang_jac2;
    float impulse, inv_lhs, rhs, rhs_wo_bias, cfm_gain, cfm_coeff;
    float lo, hi;   // impulse_bounds
    int dof;
    int kind;       // WritebackId: 0 = Dof, 1 = Limit, 2 = Motor
};

// ContactWithTwistFriction + builder (contact_with_twist_friction.rs:44-55, :601-630), one lane.
struct NormalPart {
    V3 torque_dir1, torque_dir2, ii_torque_dir1, ii_torque_dir2;
    float rhs, rhs_wo_bias, impulse, impulse_accumulator, r, cfm_factor;
};
// ContactConstraintTangentPart (contact_constraint_element.rs:14-36), one per point: FrictionModel::Coulomb only.
struct TangentPart {
    V3 torque_dir1[2], torque_dir2[2], ii_torque_dir1[2], ii_torque_dir2[2];
    float rhs[2], rhs_wo_bias[2], impulse[2], impulse_acc[2], r[3];
};
struct Constraint {
    int pair;
    bool coulomb;          // ContactWithCoulombFriction (contact_with_coulomb_friction.rs:532-549) instead of the twist form
    uint32_t id1, id2;
    int num_contacts;
    V3 dir1, tangent1;
    V3 im1, im2;
    Sdp3 ii1, ii2;
    float limit;
    NormalPart normal[MAX_MANIFOLD_POINTS];
    // tangent part
    V3 t_dp1, t_dp2, t_torque_dir1[2], t_torque_dir2[2], t_ii_torque_dir1[2], t_ii_torque_dir2[2];
    float t_rhs[2], t_rhs_wo_bias[2], t_impulse[2], t_impulse_acc[2], t_r[3];
    // twist part
    float w_rhs, w_impulse, w_impulse_acc, w_r;
    float twist_dists[MAX_MANIFOLD_POINTS];
    TangentPart tangent[MAX_MANIFOLD_POINTS];   // Coulomb: one coupled 2x2 tangent part per point
    int cids[MAX_MANIFOLD_POINTS];
    // builder
    V3 b_local_p1[MAX_MANIFOLD_POINTS], b_local_p2[MAX_MANIFOLD_POINTS];
    float b_dist[MAX_MANIFOLD_POINTS], b_restitution_seed[MAX_MANIFOLD_POINTS];
    V3 b_lfc1, b_lfc2, b_tangent_vel;
    float b_restitution;
};

struct SolverBody {
    V3 lin, ang;        // SolverVel
    Pose pose;          // SolverPose rotation+translation (CoM centred)
    Sdp3 ii;
    V3 im;
    V3 incr_lin, incr_ang;
    bool gyro;
    uint32_t flags;
};

struct World {
    Params params;
    std::vector<Body> bodies;
    std::vector<Collider> colliders;
    std::vector<Hull> hulls;           // [0] = the unit cube; ids of RB_SHAPE_CONVEX colliders (he.x)
    std::vector<Joint> joints;
    std::vector<Pair> pairs;           // sorted by (c1, c2)
    std::vector<Mask128> color_masks;  // per body (narrow_phase/mod.rs body_solver_color_masks)
    bool bp_dirty = true;
    std::vector<int> island_of;        // connected component (root body) of every dynamic body, -1 otherwise
    bool islands_dirty = true;         // the touching set or the joints changed: relabel
    std::vector<RbCollisionEvent> collision_events;      // since the last drain (EventHandler::handle_collision_event)
    std::vector<RbContactForceEvent> force_events;       // (EventHandler::handle_contact_force_event)
    std::vector<int> quarantine;       // bodies disabled because their state went non-finite (since last read)
    bool static_dirty = true;          // the sorted list of static colliders must be rebuilt
    std::vector<int> static_sorted;    // static colliders by fat min-x
    float static_max_width = 0.0f;     // widest of them along x
    // scratch
    std::vector<SolverBody> sb;
    std::vector<Constraint> cons;
    std::vector<int> order;            // constraint indices in solve order
    std::vector<int> order_color_start;  // start offset of each colour stage within `order`
    std::vector<JointRow> jrows;
    std::vector<int> jorder, jorder_color_start;
    std::vector<uint64_t> nocontact_body_pairs;  // sorted keys of joints with contacts disabled
    RbCounters counters{};
    int last_num_colors = 0;
};

int kat_solver(const char* name, const float* in, int n_in, float* out, int n_out);   // oracle_solver.cpp
int kat_world(const char* name, const float* in, int n_in, float* out, int n_out);    // oracle_world.cpp

}  // namespace orc
