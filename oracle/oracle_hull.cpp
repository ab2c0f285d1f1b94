// TEST INFRASTRUCTURE ONLY -- convex polyhedra of the CPU oracle (see oracle.h header note).
//
// What ColliderBuilder::{convex_mesh, convex_hull, round_convex_hull} (src/geometry/collider.rs:1039-1090) obtain from
// parry3d 0.30.2 (ConvexPolyhedron::from_convex_mesh, MassProperties::from_convex_polyhedron), which is not in the tree:
// face planes, the edge list, bounds, and volume / centre of mass / principal inertia of a closed convex mesh.
// Restated from the published formulas (Newell normals; signed tetrahedra against the origin with the second-moment
// identity  int x x^T dV = det/120 (a a^T + b b^T + c c^T + s s^T), s = a + b + c;  cyclic Jacobi rotations).
#include <cmath>
#include <map>
#include <utility>

#include "oracle_internal.h"

namespace orc {

namespace {
struct D3 { double x, y, z; };
inline D3 dcross(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double comp(D3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// symmetric 3x3 -> eigenvectors (columns of v) and eigenvalues, by cyclic Jacobi rotations
void diagonalise(double a[3][3], double v[3][3], double w[3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 32; ++sweep) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1.0e-30 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p) {
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    double x = a[k][p], y = a[k][q];
                    a[k][p] = c * x - s * y;
                    a[k][q] = s * x + c * y;
                }
                for (int k = 0; k < 3; ++k) {
                    double x = a[p][k], y = a[q][k];
                    a[p][k] = c * x - s * y;
                    a[q][k] = s * x + c * y;
                }
                for (int k = 0; k < 3; ++k) {
                    double x = v[k][p], y = v[k][q];
                    v[k][p] = c * x - s * y;
                    v[k][q] = s * x + c * y;
                }
            }
        }
    }
    for (int i = 0; i < 3; ++i) w[i] = a[i][i];
}

Q4 quat_of_columns(const double m[3][3]) {
    double tr = m[0][0] + m[1][1] + m[2][2], x, y, z, w;
    if (tr > 0.0) {
        double s = std::sqrt(tr + 1.0) * 2.0;
        w = 0.25 * s; x = (m[2][1] - m[1][2]) / s; y = (m[0][2] - m[2][0]) / s; z = (m[1][0] - m[0][1]) / s;
    } else if (m[0][0] > m[1][1] && m[0][0] > m[2][2]) {
        double s = std::sqrt(1.0 + m[0][0] - m[1][1] - m[2][2]) * 2.0;
        w = (m[2][1] - m[1][2]) / s; x = 0.25 * s; y = (m[0][1] + m[1][0]) / s; z = (m[0][2] + m[2][0]) / s;
    } else if (m[1][1] > m[2][2]) {
        double s = std::sqrt(1.0 + m[1][1] - m[0][0] - m[2][2]) * 2.0;
        w = (m[0][2] - m[2][0]) / s; x = (m[0][1] + m[1][0]) / s; y = 0.25 * s; z = (m[1][2] + m[2][1]) / s;
    } else {
        double s = std::sqrt(1.0 + m[2][2] - m[0][0] - m[1][1]) * 2.0;
        w = (m[1][0] - m[0][1]) / s; x = (m[0][2] + m[2][0]) / s; y = (m[1][2] + m[2][1]) / s; z = 0.25 * s;
    }
    double n = std::sqrt(x * x + y * y + z * z + w * w);
    return Q4{(float)(x / n), (float)(y / n), (float)(z / n), (float)(w / n)};
}

bool hull_mass(Hull& h) {
    double vol = 0.0;
    D3 com{0, 0, 0};
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    auto vert = [&](int i) { return D3{h.verts[i].x, h.verts[i].y, h.verts[i].z}; };
    for (size_t f = 0; f < h.face_count.size(); ++f) {
        int s = h.face_start[f], n = h.face_count[f];
        D3 a = vert(h.loops[s]);
        for (int i = 1; i + 1 < n; ++i) {
            D3 b = vert(h.loops[s + i]), c = vert(h.loops[s + i + 1]);
            double det = ddot(a, dcross(b, c));
            vol += det / 6.0;
            D3 sum{a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z};
            com.x += det * sum.x / 24.0; com.y += det * sum.y / 24.0; com.z += det * sum.z / 24.0;
            for (int r = 0; r < 3; ++r)
                for (int t = 0; t < 3; ++t)
                    C[r][t] += det / 120.0 * (comp(a, r) * comp(a, t) + comp(b, r) * comp(b, t) + comp(c, r) * comp(c, t) + comp(sum, r) * comp(sum, t));
        }
    }
    if (!(vol > 1.0e-12)) return false;
    com.x /= vol; com.y /= vol; com.z /= vol;
    double I[3][3];
    double trc = C[0][0] + C[1][1] + C[2][2], c2 = ddot(com, com);
    for (int r = 0; r < 3; ++r)
        for (int t = 0; t < 3; ++t) I[r][t] = (r == t ? trc : 0.0) - C[r][t] - vol * ((r == t ? c2 : 0.0) - comp(com, r) * comp(com, t));
    double v[3][3], w[3];
    diagonalise(I, v, w);
    D3 x0{v[0][0], v[1][0], v[2][0]}, x1{v[0][1], v[1][1], v[2][1]}, x2{v[0][2], v[1][2], v[2][2]};
    if (ddot(dcross(x0, x1), x2) < 0.0)
        for (int k = 0; k < 3; ++k) v[k][2] = -v[k][2];
    h.volume = (float)vol;
    h.com = V3{(float)com.x, (float)com.y, (float)com.z};
    h.principal_inertia = V3{(float)(w[0] > 0.0 ? w[0] : 0.0), (float)(w[1] > 0.0 ? w[1] : 0.0), (float)(w[2] > 0.0 ? w[2] : 0.0)};
    h.principal_frame = quat_of_columns(v);
    return true;
}
}  // namespace

// Tables of a closed convex mesh (either winding per face); false if it is not one within the supported limits.
bool hull_from_mesh(int nv, const float* verts, int nf, const int32_t* face_sizes, const int32_t* face_indices, Hull& h) {
    if (nv < 4 || nv > HULL_MAX_VERTS || nf < 4 || nf > HULL_MAX_FACES) return false;
    h = Hull();
    for (int i = 0; i < nv; ++i) {
        if (!std::isfinite(verts[3 * i]) || !std::isfinite(verts[3 * i + 1]) || !std::isfinite(verts[3 * i + 2])) return false;
        h.verts.push_back(V3{verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]});
    }
    D3 centroid{0, 0, 0};
    for (int i = 0; i < nv; ++i) { centroid.x += verts[3 * i] / nv; centroid.y += verts[3 * i + 1] / nv; centroid.z += verts[3 * i + 2] / nv; }
    int at = 0;
    for (int f = 0; f < nf; ++f) {
        int n = face_sizes[f];
        if (n < 3 || n > HULL_MAX_FACE_VERTS) return false;
        std::vector<int> loop(face_indices + at, face_indices + at + n);
        at += n;
        for (int v : loop)
            if (v < 0 || v >= nv) return false;
        D3 nn{0, 0, 0}, mid{0, 0, 0};
        for (int k = 0; k < n; ++k) {
            const float* p = verts + 3 * loop[k];
            const float* q = verts + 3 * loop[(k + 1) % n];
            nn.x += ((double)p[1] - q[1]) * ((double)p[2] + q[2]);
            nn.y += ((double)p[2] - q[2]) * ((double)p[0] + q[0]);
            nn.z += ((double)p[0] - q[0]) * ((double)p[1] + q[1]);
            mid.x += (double)p[0] / n; mid.y += (double)p[1] / n; mid.z += (double)p[2] / n;
        }
        double len = std::sqrt(ddot(nn, nn));
        if (!(len > 1.0e-12)) return false;
        nn.x /= len; nn.y /= len; nn.z /= len;
        double d = ddot(nn, mid);
        if (ddot(nn, centroid) - d > 0.0) {
            std::vector<int> rev(loop.rbegin(), loop.rend());
            loop = rev;
            nn = D3{-nn.x, -nn.y, -nn.z};
            d = -d;
        }
        h.face_start.push_back((int)h.loops.size());
        h.face_count.push_back(n);
        for (int v : loop) h.loops.push_back(v);
        h.normals.push_back(V3{(float)nn.x, (float)nn.y, (float)nn.z});
        h.offsets.push_back((float)d);
    }
    double size = 0.0;
    for (int i = 0; i < 3 * nv; ++i) size = std::fmax(size, (double)std::fabs(verts[i]));
    for (int f = 0; f < nf; ++f)
        for (int i = 0; i < nv; ++i) {
            double s = (double)h.normals[f].x * verts[3 * i] + (double)h.normals[f].y * verts[3 * i + 1] + (double)h.normals[f].z * verts[3 * i + 2] - h.offsets[f];
            if (s > 1.0e-4 * (1.0 + size)) return false;
        }
    std::map<std::pair<int, int>, int> directed;
    for (int f = 0; f < nf; ++f)
        for (int k = 0; k < h.face_count[f]; ++k) {
            int a = h.loops[h.face_start[f] + k], b = h.loops[h.face_start[f] + (k + 1) % h.face_count[f]];
            if (a == b || directed.count({a, b})) return false;
            directed[{a, b}] = f;
        }
    for (const auto& kv : directed) {   // ascending (v0, v1): the edge order of the separating-axis search
        int a = kv.first.first, b = kv.first.second;
        auto opp = directed.find({b, a});
        if (opp == directed.end()) return false;
        if (a < b) h.edges.push_back(HullEdge{a, b, kv.second, opp->second});
    }
    if ((int)h.edges.size() > HULL_MAX_EDGES) return false;
    h.aabb = V3{0.f, 0.f, 0.f};
    h.radius = 0.0f;
    for (const V3& p : h.verts) {
        h.aabb = V3{fmax2(h.aabb.x, std::fabs(p.x)), fmax2(h.aabb.y, std::fabs(p.y)), fmax2(h.aabb.z, std::fabs(p.z))};
        h.radius = fmax2(h.radius, std::sqrt(std::fma(p.z, p.z, std::fma(p.y, p.y, p.x * p.x))));
    }
    return hull_mass(h);
}

void hull_unit_cube(Hull& h) {
    const float v[24] = {-1, -1, -1, 1, -1, -1, 1, 1, -1, -1, 1, -1, -1, -1, 1, 1, -1, 1, 1, 1, 1, -1, 1, 1};
    const int32_t sizes[6] = {4, 4, 4, 4, 4, 4};
    const int32_t idx[24] = {0, 3, 2, 1, 4, 5, 6, 7, 0, 1, 5, 4, 2, 3, 7, 6, 1, 2, 6, 5, 0, 4, 7, 3};
    hull_from_mesh(8, v, 6, sizes, idx, h);
}

}  // namespace orc
