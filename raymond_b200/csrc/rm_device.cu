// rm_device.cu — the hot path on the GPU (sm_100a): scene flattening + upload, the
// bit-exact Scene::intersect kernels, the wavefront path tracer and the accumulator.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (rustc never contracts
// a*b+c; every f64 add/mul/div/sqrt below is one IEEE operation in the reference's order).
// Citations are relative to the reference checkout (Nyrox/raymond).

#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "rm_internal.hpp"

namespace rm {

// ===================================================================== device data model

constexpr int kMaxObjects = 64;   // objects carried in the kernel parameter block (constant bank)
constexpr int kMaxGrids = 8;

// One AccGrid resident in HBM.
//   cells : {first reference, count} per cell            (8 B, one 64-bit load per visited cell)
//   refs  : triangle indices, ascending inside a cell     (4 B per reference)
//   tri   : 96 B per triangle, 32 B aligned, 3 sectors: [v0.xyz v1.x][v1.yz v2.xy][v2.z 0 0 0]
//   nrm   : 72 B per triangle (n0, n1, n2), read once per shaded hit
struct DevGrid {
    double bmin[3], bmax[3], cell[3];
    int res[3];
    int pad;
    unsigned long long n_cells;
    const uint2* cells;
    const unsigned* refs;
    const double* tri;
    const double* nrm;
};

struct DevObject {
    int geom;       // rm::GeometryKind
    int mat;        // rm_material_kind
    int grid;       // index into DevScene::grid
    int pad;
    double g[6];    // sphere: origin xyz, radius | plane: origin xyz, normal xyz
    double color[3];// colour (Diffuse/Metal) or emitted radiance
    double rough;
};

struct DevScene {
    int n_objects;
    int n_grids;
    DevObject obj[kMaxObjects];
    DevGrid grid[kMaxGrids];
};

struct DevCamera {
    double pos[3];
    double width, height, aspect, tan_half;   // tan(fov_vert / 2 * PI / 180), evaluated once on the host (glibc)
    double focal_length, aperture_radius;
    int W, H;
    int use_dof;
    int pad;
};

// Wavefront queue entry, SoA: ray (6), throughput (3), path slot (1).
struct Queue {
    double* f[9];
    unsigned* id;
};

struct DevTotals;

struct RenderParams {
    DevCamera cam;
    unsigned long long seed;
    const unsigned* pixel_map;   // owned pixel q -> frame pixel index (y*W + x), warp = 8x4 block
    unsigned n_pixels;           // owned pixels
    unsigned first_sample, sample_stride;
    unsigned bounce_limit;
    unsigned cap;                // paths per batch (capacity of queues and of `contrib`)
    double* contrib;             // 3 planes of `cap`: radiance each path delivered
    unsigned* counts;            // counts[d] = rays queued for depth d+1
    DevTotals* totals;
};

__host__ __device__ inline unsigned stage_slot(unsigned depth) { return depth < RM_STAGE_SLOTS - 1 ? depth : RM_STAGE_SLOTS - 1; }

struct DevTotals {
    unsigned long long samples, rays, nonfinite;
    unsigned long long stage_rays[RM_STAGE_SLOTS], cells[RM_STAGE_SLOTS], tests[RM_STAGE_SLOTS], shaded[RM_STAGE_SLOTS];
};

struct WorkCount { unsigned cells, tests, shaded; };

// ===================================================================== device math (cgmath semantics)

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator-(D3 a) { return d3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return d3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ D3 operator/(D3 a, double s) { return d3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ D3 mul(D3 a, D3 b) { return d3(a.x * b.x, a.y * b.y, a.z * b.z); }
// Vector3::dot: products summed left to right
__device__ __forceinline__ double dot(D3 a, D3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// InnerSpace::normalize: multiply by the reciprocal magnitude
__device__ __forceinline__ D3 normalize(D3 a) { return a * (1.0 / sqrt(dot(a, a))); }
__device__ __forceinline__ double dist(D3 a, D3 b) { D3 d = b - a; return sqrt(dot(d, d)); }
// Matrix3::from_cols(c0, c1, c2) * v
__device__ __forceinline__ D3 mat_mul(D3 c0, D3 c1, D3 c2, D3 v) { return (c0 * v.x + c1 * v.y) + c2 * v.z; }
__device__ __forceinline__ D3 ld3(const double* p) { return d3(p[0], p[1], p[2]); }

__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// ===================================================================== intersection (bit-exact scope)

// Sphere::intersects                                       primitives/sphere.rs:11-27
__device__ __forceinline__ bool hit_sphere(const DevObject& s, D3 o, D3 d, double& t_out) {
    D3 c = ld3(s.g) - o;
    double t = dot(c, d);
    D3 q = c - t * d;
    double p = dot(q, q);
    double r2 = s.g[3] * s.g[3];
    if (p > r2) return false;
    t -= sqrt(r2 - p);
    if (t <= 0.0) return false;
    t_out = t;
    return true;
}

// Plane::intersects                                        primitives/plane.rs:11-24
__device__ __forceinline__ bool hit_plane(D3 origin, D3 normal, D3 o, D3 d, double& t_out) {
    double denom = dot(normal, -d);
    if (denom > 1e-6) {
        D3 p0l0 = origin - o;
        double t = dot(p0l0, -normal) / denom;
        if (t >= 0.0) { t_out = t; return true; }
    }
    return false;
}

// cgmath cast::<i32>(): Some(trunc) iff i32::MIN - 1 < v < i32::MAX + 1 (NaN fails)
__device__ __forceinline__ bool cast_i32(double v, int& out) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return false;
    out = __double2int_rz(v);
    return true;
}

// AccGrid::intersects = AABB::intersects + 3D-DDA + per-cell Triangle::intersects
//                                                         acc_grid.rs:89-185, aabb.rs:10-31, triangle.rs:11-44
// All quirks are kept: the z stride of the cell index is res.z (A1), the first cell holding any
// hit returns its closest hit without an in-cell check (A2), only a negative start cell is moved
// to the box entry point (A3), ties step the later axis (A4), the per-cell closest starts at
// 5712515.0 (A6).  A failed cast (the reference panics) is reported as a miss.
template <bool COUNT>
__device__ __noinline__ bool hit_grid(const DevGrid& g, D3 o, D3 d, double& t_out, unsigned& tri_out, WorkCount& wc) {
    // AABB::intersects
    double ix = 1.0 / d.x, iy = 1.0 / d.y, iz = 1.0 / d.z;
    double t1 = (g.bmin[0] - o.x) * ix, t2 = (g.bmax[0] - o.x) * ix;
    double tmin = fmin(t1, t2), tmax = fmax(t1, t2);
    t1 = (g.bmin[1] - o.y) * iy; t2 = (g.bmax[1] - o.y) * iy;
    tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2));
    t1 = (g.bmin[2] - o.z) * iz; t2 = (g.bmax[2] - o.z) * iz;
    tmin = fmax(tmin, fmin(t1, t2)); tmax = fmin(tmax, fmax(t1, t2));
    if (!(tmax > fmax(tmin, 0.0))) return false;

    const double csx = g.cell[0], csy = g.cell[1], csz = g.cell[2];
    double sx = o.x - g.bmin[0], sy = o.y - g.bmin[1], sz = o.z - g.bmin[2];
    int cx, cy, cz;
    if (!cast_i32(sx / csx, cx) || !cast_i32(sy / csy, cy) || !cast_i32(sz / csz, cz)) return false;
    if (cx < 0 || cy < 0 || cz < 0) {
        // outer_hit_position - bounding_box.min
        sx = (o.x + d.x * tmin) - g.bmin[0];
        sy = (o.y + d.y * tmin) - g.bmin[1];
        sz = (o.z + d.z * tmin) - g.bmin[2];
        if (!cast_i32(sx / csx, cx) || !cast_i32(sy / csy, cy) || !cast_i32(sz / csz, cz)) return false;
    }
    if (d.x != d.x || d.y != d.y || d.z != d.z) return false;   // signum(NaN).cast() fails
    const bool nx = d.x < 0.0, ny = d.y < 0.0, nz = d.z < 0.0;
    // f64::signum looks at the sign bit: -0.0 steps backwards
    const int stx = (__double2hiint(d.x) < 0) ? -1 : 1;
    const int sty = (__double2hiint(d.y) < 0) ? -1 : 1;
    const int stz = (__double2hiint(d.z) < 0) ? -1 : 1;
    const double tdx = (nx ? -csx : csx) / d.x;
    const double tdy = (ny ? -csy : csy) / d.y;
    const double tdz = (nz ? -csz : csz) / d.z;
    double tmx = (((double)(cx + (nx ? 0 : 1)) * csx) - sx) / d.x;
    double tmy = (((double)(cy + (ny ? 0 : 1)) * csy) - sy) / d.y;
    double tmz = (((double)(cz + (nz ? 0 : 1)) * csz) - sz) / d.z;

    const int rx = g.res[0], ry = g.res[1], rz = g.res[2];
    const unsigned long long urx = (unsigned long long)rx, urz = (unsigned long long)rz;
    for (;;) {
        unsigned long long idx = (unsigned long long)(long long)cx + urx * ((unsigned long long)(long long)cy + (unsigned long long)(long long)cz * urz);
        if (idx >= g.n_cells) return false;
        const uint2 cell = __ldg(&g.cells[idx]);
        if (COUNT) { wc.cells++; wc.tests += cell.y; }
        double closest = 5712515.0;
        bool have = false;
        unsigned best = 0;
        for (unsigned k = 0; k < cell.y; k++) {
            const unsigned ti = __ldg(&g.refs[cell.x + k]);
            const double* tp = g.tri + (size_t)ti * 12;
            double v0x, v0y, v0z, v1x, v1y, v1z, v2x, v2y, v2z, p0, p1, p2;
            ld256(tp, v0x, v0y, v0z, v1x);
            ld256(tp + 4, v1y, v1z, v2x, v2y);
            ld256(tp + 8, v2z, p0, p1, p2);
            // Triangle::intersects
            const D3 e1 = d3(v1x - v0x, v1y - v0y, v1z - v0z);
            const D3 e2 = d3(v2x - v0x, v2y - v0y, v2z - v0z);
            const D3 h = cross(d, e2);
            const double a = dot(e1, h);
            if (a < 0.00000001 && a > -0.00000001) continue;
            const double f = 1.0 / a;
            const D3 s = d3(o.x - v0x, o.y - v0y, o.z - v0z);
            const double u = f * dot(s, h);
            if (u < 0.0 || u > 1.0) continue;
            const D3 q = cross(s, e1);
            const double v = f * dot(d, q);
            if (v < 0.0 || u + v > 1.0) continue;
            const double t = f * dot(e2, q);
            if (t > 0.00000001 && t < closest) { closest = t; best = ti; have = true; }
        }
        if (have) { t_out = closest; tri_out = best; return true; }

        if (tmx < tmy) {
            if (tmx < tmz) { cx += stx; if (cx >= rx || cx < 0) return false; tmx += tdx; }
            else           { cz += stz; if (cz >= rz || cz < 0) return false; tmz += tdz; }
        } else {
            if (tmy < tmz) { cy += sty; if (cy >= ry || cy < 0) return false; tmy += tdy; }
            else           { cz += stz; if (cz >= rz || cz < 0) return false; tmz += tdz; }
        }
    }
}

struct HitRec { double t; int obj; unsigned sub; };

// Scene::intersect: first object wins ties (strict <)      scene.rs:54-74
template <bool COUNT>
__device__ __forceinline__ void scene_intersect(const DevScene& sc, D3 o, D3 d, HitRec& h, WorkCount& wc) {
    double closest = DBL_MAX;   // F_MAX
    h.obj = -1; h.sub = 0; h.t = 0.0;
    const int n = sc.n_objects;
    for (int i = 0; i < n; i++) {
        const DevObject& ob = sc.obj[i];
        double t; unsigned sub = 0; bool got;
        if (ob.geom == GEOM_PLANE) got = hit_plane(ld3(ob.g), ld3(ob.g + 3), o, d, t);
        else if (ob.geom == GEOM_SPHERE) got = hit_sphere(ob, o, d, t);
        else got = hit_grid<COUNT>(sc.grid[ob.grid], o, d, t, sub, wc);
        if (got && t < closest) { closest = t; h.t = t; h.obj = i; h.sub = sub; }
    }
}

// ===================================================================== counter-based RNG
// Philox4x32-10, key = seed, counter = (pixel, sample, depth, draw >> 1).  Draw `i` of
// (pixel, sample, depth) stands in for the i-th rand::random::<f64>() the reference makes there.

__device__ __forceinline__ void philox(unsigned long long seed, unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned w[4]) {
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const unsigned h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}
// 52 bits -> (k + 0.5) * 2^-52, strictly inside (0, 1)
__device__ __forceinline__ double u52(unsigned hi, unsigned lo) {
    const unsigned long long bits = ((unsigned long long)hi << 32) | lo;
    return (__ull2double_rn(bits >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}

// ===================================================================== camera

// generate_primary_ray                                     src/trace.rs:322-333
// jx, jy are the jitter terms (rand - 0.5).
__device__ __forceinline__ D3 primary_direction(const DevCamera& c, unsigned xi, unsigned yi, double jx, double jy) {
    const double x = (double)xi + jx;
    const double y = (double)yi + jy;
    const double px = (2.0 * ((x + 0.5) / c.width) - 1.0) * c.tan_half * c.aspect;
    const double py = (1.0 - 2.0 * ((y + 0.5) / c.height)) * c.tan_half;
    return normalize(d3(px, py, 1.0));
}

// generate_primary_ray / generate_primary_ray_with_dof     src/trace.rs:322-360
__device__ __forceinline__ void camera_ray(const DevCamera& c, unsigned long long seed, unsigned pixel, unsigned sample, D3& o, D3& d) {
    unsigned w[4];
    philox(seed, pixel, sample, 0u, 0u, w);
    const unsigned xi = pixel % (unsigned)c.W, yi = pixel / (unsigned)c.W;
    const D3 cam = ld3(c.pos);
    D3 dir = primary_direction(c, xi, yi, u52(w[0], w[1]) - 0.5, u52(w[2], w[3]) - 0.5);
    if (!c.use_dof) { o = cam; d = dir; return; }
    // rejection-sample the aperture disk (world XY at the camera's z)
    D3 start;
    for (unsigned j = 1;; j++) {
        philox(seed, pixel, sample, 0u, j, w);
        const double r1 = u52(w[0], w[1]) * 2.0 - 1.0;
        const double r2 = u52(w[2], w[3]) * 2.0 - 1.0;
        start = d3(cam.x + r1 * c.aperture_radius, cam.y + r2 * c.aperture_radius, cam.z);
        if (dist(start, cam) < c.aperture_radius) break;
    }
    // focal plane: origin = cam + (0,0,1)*focal_length, normal (0,0,-1); Plane::intersects(primary).unwrap()
    const D3 fo = cam + d3(0.0, 0.0, 1.0) * c.focal_length;
    double t = 0.0;
    hit_plane(fo, d3(0.0, 0.0, -1.0), cam, dir, t);
    const D3 end = cam + t * dir;
    o = start;
    d = normalize(end - start);
}

// ===================================================================== shading (statistical scope)

// create_coordinate_system_of_n                            src/trace.rs:408-416
__device__ __forceinline__ void onb(D3 n, D3& t, D3& b) {
    const double sign = n.z > 0.0 ? 1.0 : -1.0;
    const double a = -1.0 / (sign + n.z);
    const double bb = n.x * n.y * a;
    t = d3(1.0 + sign * n.x * n.x * a, sign * bb, -sign * n.x);
    b = d3(bb, sign + n.y * n.y * a, -n.y);
}

__device__ __forceinline__ double pow5(double x) { const double x2 = x * x; return x2 * x2 * x; }

// Triangle::get_surface_properties (Heron-area barycentrics)    triangle.rs:47-68
__device__ __forceinline__ double heron(D3 a, D3 b, D3 c) {
    const double ab = dist(a, b), ac = dist(a, c), bc = dist(b, c);
    const double s = (ab + ac + bc) / 2.0;
    return sqrt(s * (s - ab) * (s - ac) * (s - bc));
}
__device__ __forceinline__ D3 triangle_normal(const DevGrid& g, unsigned ti, D3 p) {
    const double* tp = g.tri + (size_t)ti * 12;
    const D3 v0 = ld3(tp), v1 = ld3(tp + 3), v2 = ld3(tp + 6);
    const double* np = g.nrm + (size_t)ti * 9;
    const D3 n0 = ld3(np), n1 = ld3(np + 3), n2 = ld3(np + 6);
    const double abc = heron(v0, v1, v2);
    const double abp = heron(v0, v1, p);
    const double bcp = heron(v0, v2, p);
    const double ba = abp / abc, bb = bcp / abc;
    const double bc = 1.0 - (ba + bb);
    return normalize((n2 * ba) + (n1 * bb) + (n0 * bc));
}

// One bounce of `trace` (src/trace.rs:232-320) in throughput form: the recursion multiplies the
// child radiance by a weight known before recursing, so a path's value is (prod of weights) (*)
// emission.  Returns true when the path continues with (o, d, T) updated; otherwise `result` is
// the radiance the path delivers.
__device__ __forceinline__ bool shade(const DevScene& sc, const RenderParams& rp, const HitRec& h, unsigned pixel, unsigned sample,
                                      unsigned depth, D3& o, D3& d, D3& T, D3& result, WorkCount& wc) {
    result = d3(0.0, 0.0, 0.0);
    if (h.obj < 0) return false;                                            // :242
    const DevObject& ob = sc.obj[h.obj];
    const D3 frag = o + d * h.t;                                            // :246
    if (ob.mat == RM_MATERIAL_EMISSION) { result = mul(T, ld3(ob.color)); return false; }   // :250
    if (depth >= rp.bounce_limit) return false;                             // the child call returns 0 (:235-237)
    D3 normal;                                                              // :244
    if (ob.geom == GEOM_PLANE) normal = ld3(ob.g + 3);
    else if (ob.geom == GEOM_SPHERE) normal = normalize(frag - ld3(ob.g));
    else { normal = triangle_normal(sc.grid[ob.grid], h.sub, frag); wc.shaded++; }

    const D3 color = ld3(ob.color);
    const double rough = ob.rough;
    const double metal = ob.mat == RM_MATERIAL_METAL ? 1.0 : 0.0;
    const D3 view = normalize(ld3(rp.cam.pos) - frag);                      // :256 (always the camera position)
    const D3 f0 = d3(0.04 + metal * (color.x - 0.04), 0.04 + metal * (color.y - 0.04), 0.04 + metal * (color.z - 0.04));
    unsigned w[4];
    philox(rp.seed, pixel, sample, depth, 0u, w);
    const double r = u52(w[0], w[1]);                                       // :260
    const double ra = u52(w[2], w[3]);
    philox(rp.seed, pixel, sample, depth, 1u, w);
    const double rb = u52(w[0], w[1]);
    const double prob_d = 0.5 + metal * (0.0 - 0.5);                        // :263
    D3 wgt, dir;
    double eps;
    if (r < prob_d) {
        // cosine-weighted hemisphere: theta = acos(sqrt(r1)), pdf = sqrt(r1)      :396-406
        const double ct = sqrt(ra), st = sqrt(1.0 - ra);
        double sp, cp;
        sincos(2.0 * 3.14159265358979323846 * rb, &sp, &cp);
        D3 tg, bt;
        onb(normal, tg, bt);
        dir = normalize(mat_mul(tg, normal, bt, d3(st * cp, ct, st * sp)));    // :266
        const double cos_theta = fmax(dot(normal, dir), 0.0);                   // :275
        const D3 half = normalize(dir + view);
        const double fr = pow5(1.0 - fmax(dot(half, view), 0.0));               // :277
        const D3 fres = f0 + (d3(1.0, 1.0, 1.0) - f0) * fr;
        const D3 diff = (d3(1.0, 1.0, 1.0) - fres) * (1.0 - metal);
        wgt = (mul(diff, color) * cos_theta) / (prob_d * ct);                   // :281-282
        eps = 0.00001;                                                          // :269
    } else {
        const D3 refl = normalize(-view - 2.0 * (-dot(view, normal) * normal)); // :285
        const double a = rough * rough;
        const double phi = 2.0 * 3.14159265358979323846 * ra;
        const double theta = a * sqrt(rb / (1.0 - rb));                         // :291 (used as an angle)
        double sth, cth, sp, cp;
        sincos(theta, &sth, &cth);
        sincos(phi, &sp, &cp);
        D3 tg, bt;
        onb(refl, tg, bt);
        dir = normalize(mat_mul(tg, refl, bt, d3(sth * cp, cth, sth * sp)));    // :295
        const double cos_theta = dot(normal, dir);                              // :306
        const D3 light = normalize(dir);
        const D3 half = normalize(light + view);
        const double hv = dot(half, view);
        const D3 F = f0 + (d3(1.0, 1.0, 1.0) - f0) * pow5(1.0 - hv);            // :309
        const double a2 = rough * rough;                                        // ggx_distribution :362-370
        const double nh = dot(normal, half);
        double den = (nh * nh) * (a2 - 1.0) + 1.0;
        den = fmax(3.14159265358979323846 * den * den, 1e-7);
        const double D = a2 / den;
        const double k = (rough * rough) / 8.0;                                 // geometry_smith :372-382
        const double nv = fmax(dot(normal, view), 0.0), nl = fmax(dot(normal, dir), 0.0);
        const double G = (nv / (nv * (1.0 - k) + k)) * (nl / (nl * (1.0 - k) + k));
        const D3 nom = (D * G) * F;
        const double denom = 4.0 * dot(normal, view) * cos_theta + 0.001;       // :313
        const double pdf = (D * nh) / (4.0 * hv) + 0.0001;                      // :317
        wgt = (((nom / denom) * cos_theta) / (1.0 - prob_d)) / pdf;
        eps = 0.0001;                                                           // :300
    }
    T = mul(T, wgt);
    if (T.x == 0.0 && T.y == 0.0 && T.z == 0.0) return false;   // nothing downstream can change a zero path value
    o = frag + normal * eps;
    d = dir;
    return true;
}

// ===================================================================== kernels

constexpr int kBlock = 256;

// Stage `depth` of the wavefront: [camera ray |  queue entry] -> Scene::intersect -> shade ->
// [deliver radiance | compacted queue entry for depth + 1].
template <bool FIRST, bool COUNT>
__global__ void __launch_bounds__(kBlock) k_bounce(const __grid_constant__ DevScene sc, const __grid_constant__ RenderParams rp,
                                                    const Queue qin, const Queue qout, const unsigned depth, const unsigned n_first) {
    const unsigned n = FIRST ? n_first : rp.counts[depth - 1];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warps_total = (gridDim.x * blockDim.x) >> 5;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    WorkCount wc{0u, 0u, 0u};
    for (unsigned base = warp * 32u; base < n; base += warps_total * 32u) {
        const unsigned i = base + lane;
        const bool active = i < n;
        bool cont = false;
        D3 o, d, T;
        unsigned slot = 0;
        if (active) {
            if (FIRST) {
                slot = i;
            } else {
                slot = qin.id[i];
                o = d3(qin.f[0][i], qin.f[1][i], qin.f[2][i]);
                d = d3(qin.f[3][i], qin.f[4][i], qin.f[5][i]);
                T = d3(qin.f[6][i], qin.f[7][i], qin.f[8][i]);
            }
            const unsigned q = slot % rp.n_pixels, s_local = slot / rp.n_pixels;
            const unsigned pixel = rp.pixel_map[q];
            const unsigned sample = rp.first_sample + s_local * rp.sample_stride;
            if (FIRST) {
                camera_ray(rp.cam, rp.seed, pixel, sample, o, d);
                T = d3(1.0, 1.0, 1.0);
            }
            HitRec h;
            scene_intersect<COUNT>(sc, o, d, h, wc);
            D3 result;
            cont = shade(sc, rp, h, pixel, sample, depth, o, d, T, result, wc);
            if (!cont) {
                rp.contrib[slot] = result.x;
                rp.contrib[(size_t)rp.cap + slot] = result.y;
                rp.contrib[2 * (size_t)rp.cap + slot] = result.z;
            }
        }
        // warp-aggregated compaction: one atomic per warp, survivors packed by lane rank
        const unsigned mask = __ballot_sync(0xffffffffu, cont);
        if (mask) {
            unsigned pos = 0;
            if (lane == 0) pos = atomicAdd(&rp.counts[depth], (unsigned)__popc(mask));
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(mask & ((1u << lane) - 1u));
            if (cont) {
                qout.id[pos] = slot;
                qout.f[0][pos] = o.x; qout.f[1][pos] = o.y; qout.f[2][pos] = o.z;
                qout.f[3][pos] = d.x; qout.f[4][pos] = d.y; qout.f[5][pos] = d.z;
                qout.f[6][pos] = T.x; qout.f[7][pos] = T.y; qout.f[8][pos] = T.z;
            }
        }
    }
    if (COUNT) {
        const unsigned c = __reduce_add_sync(0xffffffffu, wc.cells), t = __reduce_add_sync(0xffffffffu, wc.tests),
                       sh = __reduce_add_sync(0xffffffffu, wc.shaded);
        if (lane == 0) {
            const unsigned slot = stage_slot(depth);
            atomicAdd(&rp.totals->cells[slot], (unsigned long long)c);
            atomicAdd(&rp.totals->tests[slot], (unsigned long long)t);
            atomicAdd(&rp.totals->shaded[slot], (unsigned long long)sh);
        }
    }
}

// tile.data[..] += sample, one pass after the other          src/trace.rs:203
// Adds the batch's per-path radiance to the running sums in global sample order, so the
// association is the reference's ((0 + s0) + s1) + ...  Non-finite samples are dropped and counted
// unless RM_FLAG_KEEP_NONFINITE.
__global__ void __launch_bounds__(kBlock) k_accumulate(const __grid_constant__ RenderParams rp, double* __restrict__ accum,
                                                        const unsigned n_batch_samples, const unsigned keep_nonfinite,
                                                        DevTotals* __restrict__ totals) {
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bad = 0;
    if (q < rp.n_pixels) {
        const size_t p = (size_t)rp.pixel_map[q] * 3;
        double sx = accum[p], sy = accum[p + 1], sz = accum[p + 2];
        for (unsigned s = 0; s < n_batch_samples; s++) {
            const size_t k = (size_t)s * rp.n_pixels + q;
            const double cx = rp.contrib[k], cy = rp.contrib[(size_t)rp.cap + k], cz = rp.contrib[2 * (size_t)rp.cap + k];
            const bool finite = isfinite(cx) && isfinite(cy) && isfinite(cz);
            if (!finite) bad++;
            if (finite || keep_nonfinite) { sx += cx; sy += cy; sz += cz; }
        }
        accum[p] = sx; accum[p + 1] = sy; accum[p + 2] = sz;
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31u) == 0 && bad) atomicAdd(&totals->nonfinite, (unsigned long long)bad);
    if (q == 0) {
        unsigned long long rays = (unsigned long long)n_batch_samples * rp.n_pixels;
        if (rp.bounce_limit) totals->stage_rays[1] += rays;
        for (unsigned dpt = 1; dpt < rp.bounce_limit; dpt++) { rays += rp.counts[dpt]; totals->stage_rays[stage_slot(dpt + 1)] += rp.counts[dpt]; }
        totals->rays += rp.bounce_limit ? rays : 0ull;
        totals->samples += (unsigned long long)n_batch_samples * rp.n_pixels;
    }
}

// Scene::intersect over explicit rays (AoS rm_ray in, object / triangle index / distance out).
__global__ void __launch_bounds__(kBlock) k_hit_query(const __grid_constant__ DevScene sc, const rm_ray* __restrict__ rays, const size_t n,
                                                       long long* __restrict__ obj, unsigned long long* __restrict__ sub, double* __restrict__ dist_out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double* r = reinterpret_cast<const double*>(rays + i);
        HitRec h;
        WorkCount wc{0u, 0u, 0u};
        scene_intersect<false>(sc, d3(r[0], r[1], r[2]), d3(r[3], r[4], r[5]), h, wc);
        if (obj) obj[i] = h.obj;
        if (sub) sub[i] = h.sub;
        if (dist_out && h.obj >= 0) dist_out[i] = h.t;
    }
}

// generate_primary_ray with the jitter term forced to 0, whole frame, row-major.
__global__ void __launch_bounds__(kBlock) k_primary_rays(const __grid_constant__ DevCamera cam, rm_ray* __restrict__ rays) {
    const size_t n = (size_t)cam.W * cam.H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const D3 d = primary_direction(cam, (unsigned)(i % cam.W), (unsigned)(i / cam.W), 0.0, 0.0);
        double* r = reinterpret_cast<double*>(rays + i);
        r[0] = cam.pos[0]; r[1] = cam.pos[1]; r[2] = cam.pos[2];
        r[3] = d.x; r[4] = d.y; r[5] = d.z;
    }
}

// ===================================================================== host: CUDA plumbing

#define RM_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(RM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
    } while (0)

static int sm_count(int device) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n > 0 ? n : 148;
}

static DevCamera make_camera(const rm_camera_settings& c) {
    DevCamera d{};
    d.pos[0] = c.position.x; d.pos[1] = c.position.y; d.pos[2] = c.position.z;
    d.width = (double)c.backbuffer_width;
    d.height = (double)c.backbuffer_height;
    d.aspect = d.width / d.height;
    d.tan_half = std::tan(c.fov_vert / 2.0 * 3.14159265358979323846 / 180.0);   // src/trace.rs:329-330
    d.focal_length = c.focal_length;
    d.aperture_radius = c.aperture_radius;
    d.W = (int)c.backbuffer_width;
    d.H = (int)c.backbuffer_height;
    d.use_dof = c.aperture_radius > 0.0 ? 1 : 0;   // the reference's DoF loop never ends for radius 0
    return d;
}

}  // namespace rm

using namespace rm;

// --------------------------------------------------------------------- device scene

struct rm_device_scene {
    int device = 0;
    DevScene scene{};
    std::vector<void*> allocations;
    std::vector<std::shared_ptr<Grid>> keep;
    double upload_ms = 0.0;
    size_t bytes = 0;
    ~rm_device_scene() {
        cudaSetDevice(device);
        for (void* p : allocations) cudaFree(p);
    }
};

namespace rm {

// Page-locked staging buffer: the flattened arrays are written straight into pinned memory and
// copied H2D from there.
template <typename T>
struct Pinned {
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        n = count;
        RM_CUDA(cudaMallocHost((void**)&p, std::max<size_t>(count * sizeof(T), 32)));
        return RM_OK;
    }
    ~Pinned() { if (p) cudaFreeHost(p); }
};

template <typename T>
static int upload(rm_device_scene* ds, const Pinned<T>& host, const T** out) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(host.n * sizeof(T), 32);
    RM_CUDA(cudaMalloc(&p, bytes));
    ds->allocations.push_back(p);
    ds->bytes += bytes;
    if (host.n) RM_CUDA(cudaMemcpyAsync(p, host.p, host.n * sizeof(T), cudaMemcpyHostToDevice, 0));
    *out = (const T*)p;
    return RM_OK;
}

static int upload_grid(rm_device_scene* ds, const Grid& g, DevGrid* out) {
    DevGrid d{};
    d.bmin[0] = g.bounds.min.x; d.bmin[1] = g.bounds.min.y; d.bmin[2] = g.bounds.min.z;
    d.bmax[0] = g.bounds.max.x; d.bmax[1] = g.bounds.max.y; d.bmax[2] = g.bounds.max.z;
    d.cell[0] = g.cell_size.x; d.cell[1] = g.cell_size.y; d.cell[2] = g.cell_size.z;
    for (int a = 0; a < 3; a++) d.res[a] = (int)g.resolution[a];
    d.n_cells = g.n_cells();
    const size_t nc = (size_t)d.n_cells, nt = g.triangles.size();
    Pinned<uint2> cells;
    Pinned<unsigned> refs;
    Pinned<double> tri, nrm;
    if (int st = cells.alloc(nc)) return st;
    if (int st = refs.alloc(g.references.size())) return st;
    if (int st = tri.alloc(nt * 12)) return st;
    if (int st = nrm.alloc(nt * 9)) return st;
    for (size_t c = 0; c < nc; c++) cells.p[c] = make_uint2(g.cell_start[c], g.cell_start[c + 1] - g.cell_start[c]);
    if (!g.references.empty()) memcpy(refs.p, g.references.data(), g.references.size() * sizeof(unsigned));
    for (size_t i = 0; i < nt; i++) {
        const rm_triangle& t = g.triangles[i];
        const rm_vertex* v[3] = {&t.v0, &t.v1, &t.v2};
        for (int k = 0; k < 3; k++) {
            tri.p[i * 12 + 3 * k + 0] = v[k]->position.x; tri.p[i * 12 + 3 * k + 1] = v[k]->position.y; tri.p[i * 12 + 3 * k + 2] = v[k]->position.z;
            nrm.p[i * 9 + 3 * k + 0] = v[k]->normal.x; nrm.p[i * 9 + 3 * k + 1] = v[k]->normal.y; nrm.p[i * 9 + 3 * k + 2] = v[k]->normal.z;
        }
        tri.p[i * 12 + 9] = tri.p[i * 12 + 10] = tri.p[i * 12 + 11] = 0.0;
    }
    if (int st = upload(ds, cells, &d.cells)) return st;
    if (int st = upload(ds, refs, &d.refs)) return st;
    if (int st = upload(ds, tri, &d.tri)) return st;
    if (int st = upload(ds, nrm, &d.nrm)) return st;
    RM_CUDA(cudaStreamSynchronize(0));   // the pinned staging buffers are freed on return
    *out = d;
    return RM_OK;
}

static int build_device_scene(rm_device_scene* ds, const rm_scene* scene) {
    if (scene->objects.size() > (size_t)kMaxObjects)
        return fail(RM_ERR_UNSUPPORTED, "scene has more than " + std::to_string(kMaxObjects) + " objects");
    RM_CUDA(cudaSetDevice(ds->device));
    cudaEvent_t e0, e1;
    RM_CUDA(cudaEventCreate(&e0));
    RM_CUDA(cudaEventCreate(&e1));
    RM_CUDA(cudaEventRecord(e0, 0));
    std::map<const Grid*, int> grid_index;
    DevScene& s = ds->scene;
    s.n_objects = (int)scene->objects.size();
    s.n_grids = 0;
    for (size_t i = 0; i < scene->objects.size(); i++) {
        const Object& o = scene->objects[i];
        DevObject& d = s.obj[i];
        d.geom = o.geometry;
        d.mat = (int)o.material.kind;
        d.grid = 0;
        d.g[0] = o.origin.x; d.g[1] = o.origin.y; d.g[2] = o.origin.z;
        if (o.geometry == GEOM_SPHERE) { d.g[3] = o.radius; d.g[4] = d.g[5] = 0.0; }
        else { d.g[3] = o.normal.x; d.g[4] = o.normal.y; d.g[5] = o.normal.z; }
        d.color[0] = o.material.a.x; d.color[1] = o.material.a.y; d.color[2] = o.material.a.z;
        d.rough = o.material.p0;
        if (o.geometry == GEOM_GRID) {
            auto it = grid_index.find(o.grid.get());
            if (it == grid_index.end()) {
                if (s.n_grids >= kMaxGrids) return fail(RM_ERR_UNSUPPORTED, "scene has more than " + std::to_string(kMaxGrids) + " distinct grids");
                if (int st = upload_grid(ds, *o.grid, &s.grid[s.n_grids])) return st;
                ds->keep.push_back(o.grid);
                it = grid_index.emplace(o.grid.get(), s.n_grids++).first;
            }
            d.grid = it->second;
        }
    }
    RM_CUDA(cudaEventRecord(e1, 0));
    RM_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ds->upload_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return RM_OK;
}

}  // namespace rm

// --------------------------------------------------------------------- renderer

struct rm_renderer {
    rm_device_scene* ds = nullptr;
    bool owns_scene = false;
    rm_settings settings{};
    rm_gpu_options opt{};
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    RenderParams rp{};
    Queue q[2]{};
    double* queue_mem = nullptr;
    unsigned* id_mem = nullptr;
    unsigned* pixel_map = nullptr;
    double* accum = nullptr;
    bool owns_accum = false;
    DevTotals* totals = nullptr;
    size_t batch_spp = 1;
    uint64_t launches = 0;
    double device_ms = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    struct StageEvent { cudaEvent_t a, b; unsigned slot; };
    std::vector<StageEvent> stage_pending;
    std::vector<cudaEvent_t> event_pool;
    double stage_ms[RM_STAGE_SLOTS] = {};
    uint64_t stage_launches[RM_STAGE_SLOTS] = {};

    ~rm_renderer() {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        for (auto& p : pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
        for (auto& e : stage_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        for (auto& e : event_pool) cudaEventDestroy(e);
        cudaFree(queue_mem); cudaFree(id_mem); cudaFree(pixel_map); cudaFree(rp.contrib); cudaFree(rp.counts); cudaFree(totals);
        if (owns_accum) cudaFree(accum);
        if (owns_stream && stream) cudaStreamDestroy(stream);
        if (owns_scene) delete ds;
    }
};

namespace rm {

// Owned pixels in launch order: tiles in the reference's queue order (src/trace.rs:146-172),
// restricted to this rank's share; inside a tile, 8x4 pixel blocks so that a warp covers a compact
// screen patch (coherent primary rays walk the same grid cells).
static std::vector<unsigned> build_pixel_map(const rm_settings& s, const rm_gpu_options& o) {
    const size_t W = s.camera_settings.backbuffer_width, H = s.camera_settings.backbuffer_height;
    std::vector<TileRect> tiles = tile_layout(W, H, s.tile_size[0], s.tile_size[1]);
    const int world = o.world_size > 1 ? o.world_size : 1;
    std::vector<unsigned> map;
    map.reserve(W * H / (o.partition == RM_PARTITION_TILES ? world : 1) + 1024);
    for (size_t t = 0; t < tiles.size(); t++) {
        if (o.partition == RM_PARTITION_TILES && (int)(t % (size_t)world) != o.rank) continue;
        const TileRect& r = tiles[t];
        for (size_t by = 0; by < r.height; by += 4)
            for (size_t bx = 0; bx < r.width; bx += 8)
                for (size_t y = by; y < std::min(by + 4, r.height); y++)
                    for (size_t x = bx; x < std::min(bx + 8, r.width); x++)
                        map.push_back((unsigned)((r.left + x) + (r.top + y) * W));
    }
    return map;
}

static int renderer_init(rm_renderer* r) {
    const rm_settings& s = r->settings;
    const size_t W = s.camera_settings.backbuffer_width, H = s.camera_settings.backbuffer_height;
    if (W == 0 || H == 0 || s.tile_size[0] == 0 || s.tile_size[1] == 0) return fail(RM_ERR_INVALID_ARGUMENT, "empty backbuffer or tile size");
    if (W * H >= 0xffffffffull) return fail(RM_ERR_UNSUPPORTED, "backbuffer has 2^32 or more pixels");
    if (s.bounce_limit > 1u << 20) return fail(RM_ERR_UNSUPPORTED, "bounce_limit too large");
    if (r->opt.world_size > 1 && (r->opt.rank < 0 || r->opt.rank >= r->opt.world_size)) return fail(RM_ERR_INVALID_ARGUMENT, "rank outside world_size");
    RM_CUDA(cudaSetDevice(r->device));
    r->sms = sm_count(r->device);
    if (r->opt.stream) r->stream = (cudaStream_t)r->opt.stream;
    else { RM_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking)); r->owns_stream = true; }

    std::vector<unsigned> map = build_pixel_map(s, r->opt);
    const size_t npix = map.size();
    RM_CUDA(cudaMalloc(&r->pixel_map, std::max<size_t>(npix, 1) * sizeof(unsigned)));
    if (npix) RM_CUDA(cudaMemcpyAsync(r->pixel_map, map.data(), npix * sizeof(unsigned), cudaMemcpyHostToDevice, r->stream));
    RM_CUDA(cudaStreamSynchronize(r->stream));

    // batch: enough paths in flight to fill the machine many times over, bounded in memory
    size_t spp = r->opt.batch_spp;
    if (spp == 0) {
        const size_t target = (size_t)8 << 20;   // ~8 Mi paths per wavefront batch
        spp = npix ? std::max<size_t>(1, target / npix) : 1;
        spp = std::min<size_t>(spp, 64);
    }
    while (spp > 1 && spp * npix >= 0xffffffffull) spp--;
    r->batch_spp = spp;
    const size_t cap = std::max<size_t>(npix * spp, 32);

    RM_CUDA(cudaMalloc(&r->queue_mem, cap * 9 * 2 * sizeof(double)));
    RM_CUDA(cudaMalloc(&r->id_mem, cap * 2 * sizeof(unsigned)));
    for (int b = 0; b < 2; b++) {
        for (int k = 0; k < 9; k++) r->q[b].f[k] = r->queue_mem + ((size_t)b * 9 + k) * cap;
        r->q[b].id = r->id_mem + (size_t)b * cap;
    }
    RenderParams& rp = r->rp;
    rp.cam = make_camera(s.camera_settings);
    rp.seed = r->opt.seed;
    rp.pixel_map = r->pixel_map;
    rp.n_pixels = (unsigned)npix;
    rp.bounce_limit = (unsigned)s.bounce_limit;
    rp.cap = (unsigned)cap;
    RM_CUDA(cudaMalloc(&rp.contrib, cap * 3 * sizeof(double)));
    RM_CUDA(cudaMalloc(&rp.counts, (s.bounce_limit + 2) * sizeof(unsigned)));
    RM_CUDA(cudaMalloc(&r->totals, sizeof(DevTotals)));
    rp.totals = r->totals;
    RM_CUDA(cudaMemsetAsync(r->totals, 0, sizeof(DevTotals), r->stream));
    if (r->opt.accum_device) r->accum = (double*)r->opt.accum_device;
    else { RM_CUDA(cudaMalloc(&r->accum, W * H * 3 * sizeof(double))); r->owns_accum = true; }
    RM_CUDA(cudaMemsetAsync(r->accum, 0, W * H * 3 * sizeof(double), r->stream));
    return RM_OK;
}

static cudaEvent_t take_event(rm_renderer* r) {
    if (!r->event_pool.empty()) { cudaEvent_t e = r->event_pool.back(); r->event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Brackets one launch with events when RM_FLAG_STAGE_TIMING is set.
struct StageTimer {
    rm_renderer* r; unsigned slot; cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(rm_renderer* r_, unsigned slot_) : r(r_), slot(slot_) {
        if (r->opt.flags & RM_FLAG_STAGE_TIMING) { a = take_event(r); b = take_event(r); cudaEventRecord(a, r->stream); }
    }
    ~StageTimer() {
        r->launches++;
        r->stage_launches[slot]++;
        if (a) { cudaEventRecord(b, r->stream); r->stage_pending.push_back({a, b, slot}); }
    }
};

static int renderer_batch(rm_renderer* r, unsigned first_sample, unsigned n_samples, unsigned stride) {
    RenderParams rp = r->rp;
    rp.first_sample = first_sample;
    rp.sample_stride = stride;
    const unsigned n_paths = rp.n_pixels * n_samples;
    const unsigned limit = rp.bounce_limit;
    if (n_paths == 0) return RM_OK;
    const unsigned persistent_grid = (unsigned)r->sms * 8u;
    RM_CUDA(cudaMemsetAsync(rp.counts, 0, (limit + 2) * sizeof(unsigned), r->stream));
    if (limit >= 1) {
        const unsigned grid = std::min<unsigned>((n_paths + kBlock - 1) / kBlock, persistent_grid * 4u);
        const bool count = (r->opt.flags & RM_FLAG_COUNT_WORK) != 0;
        {
            StageTimer st(r, stage_slot(1));
            if (count) k_bounce<true, true><<<grid, kBlock, 0, r->stream>>>(r->ds->scene, rp, r->q[1], r->q[0], 1u, n_paths);
            else k_bounce<true, false><<<grid, kBlock, 0, r->stream>>>(r->ds->scene, rp, r->q[1], r->q[0], 1u, n_paths);
        }
        for (unsigned depth = 2; depth <= limit; depth++) {
            StageTimer st(r, stage_slot(depth));
            if (count) k_bounce<false, true><<<persistent_grid, kBlock, 0, r->stream>>>(r->ds->scene, rp, r->q[depth & 1], r->q[(depth + 1) & 1], depth, 0u);
            else k_bounce<false, false><<<persistent_grid, kBlock, 0, r->stream>>>(r->ds->scene, rp, r->q[depth & 1], r->q[(depth + 1) & 1], depth, 0u);
        }
    } else {
        RM_CUDA(cudaMemsetAsync(rp.contrib, 0, (size_t)rp.cap * 3 * sizeof(double), r->stream));   // trace() returns 0 at depth 1 > 0
    }
    {
        StageTimer st(r, 0);
        k_accumulate<<<(rp.n_pixels + kBlock - 1) / kBlock, kBlock, 0, r->stream>>>(rp, r->accum, n_samples,
                                                                                   (r->opt.flags & RM_FLAG_KEEP_NONFINITE) ? 1u : 0u, r->totals);
    }
    RM_CUDA(cudaGetLastError());
    return RM_OK;
}

static void harvest_events(rm_renderer* r) {
    for (auto& p : r->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) r->device_ms += ms;
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    r->pending.clear();
    for (auto& e : r->stage_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) r->stage_ms[e.slot] += ms;
        r->event_pool.push_back(e.a);
        r->event_pool.push_back(e.b);
    }
    r->stage_pending.clear();
}

}  // namespace rm

// ===================================================================== C ABI (device level)

extern "C" {

rm_device_scene* rm_device_scene_create(const rm_scene* scene, int device) {
    if (!scene) { set_error("rm_device_scene_create: null scene"); return nullptr; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        set_error(std::string("no usable CUDA device ") + std::to_string(device) + " (" + (e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range") +
                  "); this library has no CPU path");
        return nullptr;
    }
    rm_device_scene* ds = new rm_device_scene();
    ds->device = device;
    if (build_device_scene(ds, scene) != RM_OK) { delete ds; return nullptr; }
    return ds;
}

void rm_device_scene_destroy(rm_device_scene* ds) { delete ds; }

int rm_device_scene_intersect(rm_device_scene* ds, const rm_ray* rays, size_t count, int64_t* obj, uint64_t* sub, double* distance, void* stream) {
    if (!ds || (!rays && count)) return fail(RM_ERR_INVALID_ARGUMENT, "rm_device_scene_intersect: null argument");
    if (count == 0) return RM_OK;
    RM_CUDA(cudaSetDevice(ds->device));
    const size_t blocks = std::min<size_t>((count + kBlock - 1) / kBlock, (size_t)sm_count(ds->device) * 32);
    k_hit_query<<<(unsigned)blocks, kBlock, 0, (cudaStream_t)stream>>>(ds->scene, rays, count, (long long*)obj, (unsigned long long*)sub, distance);
    RM_CUDA(cudaGetLastError());
    return RM_OK;
}

int rm_primary_rays_device(const rm_camera_settings* camera, int device, rm_ray* rays, void* stream) {
    if (!camera || !rays) return fail(RM_ERR_INVALID_ARGUMENT, "rm_primary_rays_device: null argument");
    RM_CUDA(cudaSetDevice(device));
    DevCamera cam = make_camera(*camera);
    const size_t n = (size_t)cam.W * cam.H;
    if (n == 0) return RM_OK;
    const size_t blocks = std::min<size_t>((n + kBlock - 1) / kBlock, (size_t)sm_count(device) * 32);
    k_primary_rays<<<(unsigned)blocks, kBlock, 0, (cudaStream_t)stream>>>(cam, rays);
    RM_CUDA(cudaGetLastError());
    return RM_OK;
}

int rm_scene_intersect(const rm_scene* scene, int device, const rm_ray* rays, size_t count, int64_t* obj, uint64_t* sub, double* distance) {
    if (!scene || (!rays && count)) return fail(RM_ERR_INVALID_ARGUMENT, "rm_scene_intersect: null argument");
    rm_device_scene* ds = rm_device_scene_create(scene, device);
    if (!ds) return RM_ERR_CUDA;
    int st = RM_OK;
    rm_ray* d_rays = nullptr; int64_t* d_obj = nullptr; uint64_t* d_sub = nullptr; double* d_t = nullptr;
    auto cleanup = [&]() { cudaFree(d_rays); cudaFree(d_obj); cudaFree(d_sub); cudaFree(d_t); delete ds; };
#define RM_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { st = fail(RM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); cleanup(); return st; } } while (0)
    if (count) {
        RM_TRY(cudaMalloc(&d_rays, count * sizeof(rm_ray)));
        RM_TRY(cudaMalloc(&d_obj, count * sizeof(int64_t)));
        RM_TRY(cudaMalloc(&d_sub, count * sizeof(uint64_t)));
        RM_TRY(cudaMalloc(&d_t, count * sizeof(double)));
        RM_TRY(cudaMemcpy(d_rays, rays, count * sizeof(rm_ray), cudaMemcpyHostToDevice));
        if (distance) RM_TRY(cudaMemcpy(d_t, distance, count * sizeof(double), cudaMemcpyHostToDevice));   // misses leave it untouched
        st = rm_device_scene_intersect(ds, d_rays, count, d_obj, d_sub, d_t, nullptr);
        if (st == RM_OK) {
            RM_TRY(cudaDeviceSynchronize());
            if (obj) RM_TRY(cudaMemcpy(obj, d_obj, count * sizeof(int64_t), cudaMemcpyDeviceToHost));
            if (sub) RM_TRY(cudaMemcpy(sub, d_sub, count * sizeof(uint64_t), cudaMemcpyDeviceToHost));
            if (distance) RM_TRY(cudaMemcpy(distance, d_t, count * sizeof(double), cudaMemcpyDeviceToHost));
        }
    }
#undef RM_TRY
    cleanup();
    return st;
}

rm_renderer* rm_renderer_create_on(rm_device_scene* ds, const rm_settings* settings, const rm_gpu_options* options) {
    if (!ds || !settings) { set_error("rm_renderer_create_on: null argument"); return nullptr; }
    rm_renderer* r = new rm_renderer();
    r->ds = ds;
    r->settings = *settings;
    if (options) r->opt = *options;
    r->opt.device = ds->device;
    r->device = ds->device;
    if (renderer_init(r) != RM_OK) { delete r; return nullptr; }
    return r;
}

rm_renderer* rm_renderer_create(const rm_scene* scene, const rm_settings* settings, const rm_gpu_options* options) {
    if (!scene || !settings) { set_error("rm_renderer_create: null argument"); return nullptr; }
    rm_device_scene* ds = rm_device_scene_create(scene, options ? options->device : 0);
    if (!ds) return nullptr;
    rm_renderer* r = rm_renderer_create_on(ds, settings, options);
    if (!r) { delete ds; return nullptr; }
    r->owns_scene = true;
    return r;
}

int rm_renderer_render(rm_renderer* r, size_t first_sample, size_t count, size_t stride) {
    if (!r) return fail(RM_ERR_INVALID_ARGUMENT, "rm_renderer_render: null renderer");
    if (stride == 0) stride = 1;
    if (first_sample + count * stride >= 0xffffffffull) return fail(RM_ERR_UNSUPPORTED, "sample index beyond 2^32");
    RM_CUDA(cudaSetDevice(r->device));
    cudaEvent_t e0, e1;
    RM_CUDA(cudaEventCreate(&e0));
    RM_CUDA(cudaEventCreate(&e1));
    RM_CUDA(cudaEventRecord(e0, r->stream));
    size_t done = 0;
    int st = RM_OK;
    while (done < count && st == RM_OK) {
        const size_t nb = std::min(r->batch_spp, count - done);
        st = renderer_batch(r, (unsigned)(first_sample + done * stride), (unsigned)nb, (unsigned)stride);
        done += nb;
    }
    cudaEventRecord(e1, r->stream);
    r->pending.emplace_back(e0, e1);
    return st;
}

void* rm_renderer_accum_device(rm_renderer* r) { return r ? r->accum : nullptr; }

int rm_renderer_clear(rm_renderer* r) {
    if (!r) return fail(RM_ERR_INVALID_ARGUMENT, "rm_renderer_clear: null renderer");
    RM_CUDA(cudaSetDevice(r->device));
    const size_t n = r->settings.camera_settings.backbuffer_width * r->settings.camera_settings.backbuffer_height * 3;
    RM_CUDA(cudaMemsetAsync(r->accum, 0, n * sizeof(double), r->stream));   // statistics stay cumulative
    return RM_OK;
}

int rm_renderer_sync(rm_renderer* r) {
    if (!r) return fail(RM_ERR_INVALID_ARGUMENT, "rm_renderer_sync: null renderer");
    RM_CUDA(cudaSetDevice(r->device));
    RM_CUDA(cudaStreamSynchronize(r->stream));
    harvest_events(r);
    return RM_OK;
}

int rm_renderer_read_sums(rm_renderer* r, rm_vec3* out) {
    if (!r || !out) return fail(RM_ERR_INVALID_ARGUMENT, "rm_renderer_read_sums: null argument");
    RM_CUDA(cudaSetDevice(r->device));
    const size_t n = r->settings.camera_settings.backbuffer_width * r->settings.camera_settings.backbuffer_height;
    RM_CUDA(cudaMemcpyAsync(out, r->accum, n * sizeof(rm_vec3), cudaMemcpyDeviceToHost, r->stream));
    RM_CUDA(cudaStreamSynchronize(r->stream));
    harvest_events(r);
    return RM_OK;
}

int rm_renderer_read_frame(rm_renderer* r, size_t sample_count, rm_vec3* out) {
    if (int st = rm_renderer_read_sums(r, out)) return st;
    // tile.data[..] / tile.sample_count as f64             src/trace.rs:95
    const size_t n = r->settings.camera_settings.backbuffer_width * r->settings.camera_settings.backbuffer_height;
    const double c = (double)sample_count;
    for (size_t i = 0; i < n; i++) { out[i].x /= c; out[i].y /= c; out[i].z /= c; }
    return RM_OK;
}

int rm_renderer_stats(rm_renderer* r, rm_stats* out) {
    if (!r || !out) return fail(RM_ERR_INVALID_ARGUMENT, "rm_renderer_stats: null argument");
    RM_CUDA(cudaSetDevice(r->device));
    RM_CUDA(cudaStreamSynchronize(r->stream));
    harvest_events(r);
    DevTotals t{};
    RM_CUDA(cudaMemcpy(&t, r->totals, sizeof(t), cudaMemcpyDeviceToHost));
    out->samples = t.samples;
    out->rays = t.rays;
    out->nonfinite_samples = t.nonfinite;
    out->kernel_launches = r->launches;
    out->device_ms = r->device_ms;
    out->upload_ms = r->ds->upload_ms;
    out->upload_bytes = r->ds->bytes + (uint64_t)r->rp.n_pixels * sizeof(unsigned);
    return RM_OK;
}

int rm_renderer_stage_stats(rm_renderer* r, rm_stage_stats* out) {
    if (!r || !out) return fail(RM_ERR_INVALID_ARGUMENT, "rm_renderer_stage_stats: null argument");
    RM_CUDA(cudaSetDevice(r->device));
    RM_CUDA(cudaStreamSynchronize(r->stream));
    harvest_events(r);
    DevTotals t{};
    RM_CUDA(cudaMemcpy(&t, r->totals, sizeof(t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < RM_STAGE_SLOTS; i++) {
        out->ms[i] = r->stage_ms[i];
        out->launches[i] = r->stage_launches[i];
        out->rays[i] = t.stage_rays[i];
        out->cells[i] = t.cells[i];
        out->triangle_tests[i] = t.tests[i];
        out->shaded_triangles[i] = t.shaded[i];
    }
    return RM_OK;
}

void rm_renderer_destroy(rm_renderer* r) { delete r; }

}  // extern "C"
