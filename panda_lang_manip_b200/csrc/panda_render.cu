// panda_render.cu -- analytic depth / point-cloud rendering of the primitive scenes, batched over environments.
//
// Replaces, for the bodies the kernels simulate, the fork's camera path: reference panda_gym/pybullet.py:149-264 (render:
// getCameraImage -> depth buffer + colours -> deprojection with inv(P V) -> "infinite depth" and workspace filters) with the camera
// of :70-107 (computeViewMatrixFromYawPitchRoll, computeProjectionMatrixFOV fov 60, near 0.1, far 100).  The reference rasterises
// the robot's visual meshes with OpenGL; those meshes are not part of /root/reference, so the robot is drawn as the boxes the physics
// uses (arm links: the inertia AABBs of panda_model.h; hand and fingers: the collision boxes of panda_scene.h) -- "primitives only".
//
// One thread per pixel, 16 x 16 pixel tiles; a tile culls the env's <= 16 oriented boxes / z-cylinders (built once per env by
// render_setup_kernel) to those whose screen rectangle touches it, rays are cast against that short list from shared memory and
// everything a pixel produces is written once (21-22 B per pixel).  Round-2 ncu: the un-culled version was instruction-bound
// (2,330 instructions per pixel, 4.2 ms for 256 x 480 x 480), not write-bound.
#include <cuda_runtime.h>
#include <math.h>
#include "panda_kernels.cuh"
#include "panda_render.h"

namespace pg {

extern long long g_launches;

template <typename T>
__global__ void __launch_bounds__(BLOCK) render_setup_kernel(const __grid_constant__ EnvDev<T> E, const RenderScene R, RenderPrim* __restrict__ prims) {
    const int i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= E.n) return;
    RenderPrim* P = prims + (size_t)i * RENDER_MAX_PRIMS;
    int np = 0;
    auto put = [&](int kind, int id, float cx, float cy, float cz, const float* X, const float* Y, const float* Z, float hx, float hy, float hz) {
        RenderPrim& p = P[np++];
        p.kind = kind; p.id = id; p.c[0] = cx; p.c[1] = cy; p.c[2] = cz; p.h[0] = hx; p.h[1] = hy; p.h[2] = hz;
        for (int k = 0; k < 3; k++) { p.X[k] = X[k]; p.Y[k] = Y[k]; p.Z[k] = Z[k]; }
    };
    const float ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
    if (R.has_plane) put(0, RENDER_ID_PLANE, 0.f, 0.f, R.plane_z - 0.01f, ex, ey, ez, 3.f, 3.f, 0.01f);                                   // pybullet.py:726-739
    if (R.has_table) put(0, RENDER_ID_TABLE, 0.5f * (R.table[0] + R.table[1]), 0.5f * (R.table[2] + R.table[3]), -0.5f * R.table_height, ex, ey, ez,
                         0.5f * (R.table[1] - R.table[0]), 0.5f * (R.table[3] - R.table[2]), 0.5f * R.table_height);                        // pybullet.py:741-771
    const int n = E.n;
    for (int o = 0; o < R.nobj; o++) {
        const T* p = E.obj + (size_t)o * 13 * n + i;
        Rot<T> Ro = quat_rot(p[3 * n], p[4 * n], p[5 * n], p[6 * n]);
        const float X[3] = {(float)Ro.X.x, (float)Ro.X.y, (float)Ro.X.z}, Y[3] = {(float)Ro.Y.x, (float)Ro.Y.y, (float)Ro.Y.z}, Z[3] = {(float)Ro.Z.x, (float)Ro.Z.y, (float)Ro.Z.z};
        put(E.S.shape[o] == SH_CYL ? 1 : 0, RENDER_ID_OBJECT + o, (float)p[0], (float)p[n], (float)p[2 * n], X, Y, Z, (float)E.S.half[o][0], (float)E.S.half[o][1], (float)E.S.half[o][2]);
    }
    if (R.has_robot) {
        T q[ND];
        for (int d = 0; d < ND; d++) q[d] = E.q[d * n + i];
        Frame<T> F[7]; fk_arm(E.M, q, F);
        put(0, RENDER_ID_ROBOT, (float)E.M.base[0] + R.base_c[0], (float)E.M.base[1] + R.base_c[1], (float)E.M.base[2] + R.base_c[2], ex, ey, ez, R.base_h[0], R.base_h[1], R.base_h[2]);
        for (int l = 0; l < 11; l++) {
            if (l == 7) continue;                                   // panda_link8: massless, no shape
            Frame<T> L = link_frame_from(E.M, F, q, l);
            const float* c = R.link_c[l]; const float* h = R.link_h[l];
            if (!(h[0] > 0.f)) continue;
            V3<T> w = L.p + L.X * (T)c[0] + L.Y * (T)c[1] + L.Z * (T)c[2];
            const float X[3] = {(float)L.X.x, (float)L.X.y, (float)L.X.z}, Y[3] = {(float)L.Y.x, (float)L.Y.y, (float)L.Y.z}, Z[3] = {(float)L.Z.x, (float)L.Z.y, (float)L.Z.z};
            put(0, RENDER_ID_ROBOT + 1 + l, (float)w.x, (float)w.y, (float)w.z, X, Y, Z, h[0], h[1], h[2]);
        }
    }
    for (; np < RENDER_MAX_PRIMS; np++) P[np].kind = -1;
}

// slab test in the primitive's frame; returns the entry distance (ray parameter; dir has unit forward component, so it IS the eye depth)
// lo: the eye in the primitive's frame (the same for every pixel of a tile: computed once per tile and primitive)
__device__ __forceinline__ float hit_box(const RenderPrim& p, const float* lo, const float* d, float* nrm) {
    const float ld[3] = {d[0] * p.X[0] + d[1] * p.X[1] + d[2] * p.X[2], d[0] * p.Y[0] + d[1] * p.Y[1] + d[2] * p.Y[2], d[0] * p.Z[0] + d[1] * p.Z[1] + d[2] * p.Z[2]};
    float t0 = -1e30f, t1 = 1e30f; int ax = 0; float sg = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float inv = __fdividef(1.0f, ld[k]);             // MUFU.RCP (1 ulp; the IEEE division is ~8 instructions, three per primitive and pixel); +-inf for an axis-parallel ray
        float a = (-p.h[k] - lo[k]) * inv, b = (p.h[k] - lo[k]) * inv;
        if (ld[k] == 0.f) { if (fabsf(lo[k]) > p.h[k]) return -1.f; a = -1e30f; b = 1e30f; }
        const float near = fminf(a, b), far = fmaxf(a, b);
        if (near > t0) { t0 = near; ax = k; sg = ld[k] > 0.f ? -1.f : 1.f; }
        t1 = fminf(t1, far);
    }
    if (t0 > t1 || t1 <= 0.f || t0 <= 0.f) return -1.f;         // miss, behind the camera, or the camera is inside the box
    const float* A = ax == 0 ? p.X : (ax == 1 ? p.Y : p.Z);
    nrm[0] = sg * A[0]; nrm[1] = sg * A[1]; nrm[2] = sg * A[2];
    return t0;
}
__device__ __forceinline__ float hit_cyl(const RenderPrim& p, const float* lo, const float* d, float* nrm) {       // z-cylinder: radius h[0], half height h[2]
    const float ld[3] = {d[0] * p.X[0] + d[1] * p.X[1] + d[2] * p.X[2], d[0] * p.Y[0] + d[1] * p.Y[1] + d[2] * p.Y[2], d[0] * p.Z[0] + d[1] * p.Z[1] + d[2] * p.Z[2]};
    const float r = p.h[0], hz = p.h[2];
    float best = -1.f;
    const float a = ld[0] * ld[0] + ld[1] * ld[1], b = lo[0] * ld[0] + lo[1] * ld[1], c = lo[0] * lo[0] + lo[1] * lo[1] - r * r;
    if (a > 0.f) {
        const float disc = b * b - a * c;
        if (disc >= 0.f) {
            const float t = (-b - sqrtf(disc)) / a, z = lo[2] + t * ld[2];
            if (t > 0.f && fabsf(z) <= hz) {
                best = t;
                const float nx = (lo[0] + t * ld[0]) / r, ny = (lo[1] + t * ld[1]) / r;
                nrm[0] = nx * p.X[0] + ny * p.Y[0]; nrm[1] = nx * p.X[1] + ny * p.Y[1]; nrm[2] = nx * p.X[2] + ny * p.Y[2];
            }
        }
    }
    if (ld[2] != 0.f) {
        const float sgn = ld[2] > 0.f ? -1.f : 1.f;            // the cap facing the ray
        const float t = (sgn * hz - lo[2]) / ld[2], x = lo[0] + t * ld[0], y = lo[1] + t * ld[1];
        if (t > 0.f && x * x + y * y <= r * r && (best < 0.f || t < best)) { best = t; nrm[0] = sgn * p.Z[0]; nrm[1] = sgn * p.Z[1]; nrm[2] = sgn * p.Z[2]; }
    }
    return best;
}

// One block = one 16 x 16 pixel tile of one env.  The tile first culls the env's primitive list: thread (k, c) projects corner c of
// primitive k's bounding box, a primitive stays in the tile's list when its screen rectangle (grown by a pixel) overlaps the tile or a
// corner lies at / behind the eye plane (conservative) -- a pixel then tests 1-4 primitives instead of up to 16, which is what the
// kernel's time went into (ncu, round 2: 87 % issue-active, 2,330 instructions per pixel before the culling).  Order of the list is
// kept, so ties resolve as without it.  The deprojected points (3 floats per pixel) leave through shared memory as contiguous runs.
constexpr int RT = 16;      // tile edge
__global__ void __launch_bounds__(RT * RT) render_kernel(const RenderPrim* __restrict__ prims, const RenderCamera C, float* __restrict__ depth, uchar4* __restrict__ rgba,
                                                         unsigned char* __restrict__ seg, float* __restrict__ points, unsigned char* __restrict__ valid) {
    __shared__ RenderPrim s_p[RENDER_MAX_PRIMS];
    __shared__ int s_lo[RENDER_MAX_PRIMS][2], s_hi[RENDER_MAX_PRIMS][2], s_behind[RENDER_MAX_PRIMS], s_list[RENDER_MAX_PRIMS], s_n;
    __shared__ float s_pts[RT * RT * 3], s_eye[RENDER_MAX_PRIMS][3];
    const int env = blockIdx.z, tid = threadIdx.x;
    {   // the env's primitive list: RENDER_MAX_PRIMS x 24 words, loaded cooperatively
        const int words = RENDER_MAX_PRIMS * (int)(sizeof(RenderPrim) / 4);
        const int* src = reinterpret_cast<const int*>(prims + (size_t)env * RENDER_MAX_PRIMS);
        for (int k = tid; k < words; k += RT * RT) reinterpret_cast<int*>(s_p)[k] = src[k];
        if (tid < RENDER_MAX_PRIMS) { s_lo[tid][0] = s_lo[tid][1] = 1 << 30; s_hi[tid][0] = s_hi[tid][1] = -(1 << 30); s_behind[tid] = 0; }
    }
    __syncthreads();
    if (tid >= RT * RT - RENDER_MAX_PRIMS) {      // the eye in each primitive's frame (the last 16 threads: the first 128 project corners)
        const RenderPrim& p = s_p[tid - (RT * RT - RENDER_MAX_PRIMS)];
        const float rel[3] = {C.eye[0] - p.c[0], C.eye[1] - p.c[1], C.eye[2] - p.c[2]};
        float* e = s_eye[tid - (RT * RT - RENDER_MAX_PRIMS)];
        e[0] = rel[0] * p.X[0] + rel[1] * p.X[1] + rel[2] * p.X[2]; e[1] = rel[0] * p.Y[0] + rel[1] * p.Y[1] + rel[2] * p.Y[2]; e[2] = rel[0] * p.Z[0] + rel[1] * p.Z[1] + rel[2] * p.Z[2];
    }
    if (tid < RENDER_MAX_PRIMS * 8) {   // corner c of primitive k -> pixel coordinates
        const int k = tid >> 3, c = tid & 7;
        const RenderPrim& p = s_p[k];
        if (p.kind >= 0) {
            const float sx = (c & 1) ? p.h[0] : -p.h[0], sy = (c & 2) ? p.h[1] : -p.h[1], sz = (c & 4) ? p.h[2] : -p.h[2];
            const float v[3] = {p.c[0] + sx * p.X[0] + sy * p.Y[0] + sz * p.Z[0] - C.eye[0], p.c[1] + sx * p.X[1] + sy * p.Y[1] + sz * p.Z[1] - C.eye[1],
                                p.c[2] + sx * p.X[2] + sy * p.Y[2] + sz * p.Z[2] - C.eye[2]};
            const float zc = v[0] * C.fwd[0] + v[1] * C.fwd[1] + v[2] * C.fwd[2];
            if (zc <= 1e-3f) atomicOr(&s_behind[k], 1);
            else {
                const float xc = v[0] * C.right[0] + v[1] * C.right[1] + v[2] * C.right[2], yc = v[0] * C.up[0] + v[1] * C.up[1] + v[2] * C.up[2];
                const float xn = xc / (zc * C.tan_half_fov * C.aspect), yn = yc / (zc * C.tan_half_fov);
                const float col = fminf(fmaxf((xn + 1.0f) * 0.5f * C.width, -1e6f), 1e6f), row = fminf(fmaxf((1.0f - yn) * 0.5f * C.height, -1e6f), 1e6f);
                atomicMin(&s_lo[k][0], (int)floorf(col) - 1); atomicMax(&s_hi[k][0], (int)ceilf(col) + 1);
                atomicMin(&s_lo[k][1], (int)floorf(row) - 1); atomicMax(&s_hi[k][1], (int)ceilf(row) + 1);
            }
        }
    }
    __syncthreads();
    const int x0 = blockIdx.x * RT, y0 = blockIdx.y * RT;
    if (tid < 32) {     // ordered compaction by the first warp: lane k decides primitive k
        bool in = false;
        if (tid < RENDER_MAX_PRIMS && s_p[tid].kind >= 0)
            in = s_behind[tid] || (s_lo[tid][0] <= x0 + RT - 1 && s_hi[tid][0] >= x0 && s_lo[tid][1] <= y0 + RT - 1 && s_hi[tid][1] >= y0);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (in) s_list[__popc(m & ((1u << tid) - 1u))] = tid;
        if (tid == 0) s_n = __popc(m);
    }
    __syncthreads();
    const int col = x0 + (tid & (RT - 1)), row = y0 + (tid >> 4), npix = C.width * C.height;
    const bool inside = col < C.width && row < C.height;
    float best = 1e30f, bn[3] = {0.f, 0.f, 1.f}; int id = 0;
    // ray through the pixel centre; dir = forward + x right + y up, unit forward component
    const float xn = (col + 0.5f) * (2.0f / C.width) - 1.0f, yn = 1.0f - (row + 0.5f) * (2.0f / C.height);
    const float dx = xn * C.tan_half_fov * C.aspect, dy = yn * C.tan_half_fov;
    const float d[3] = {C.fwd[0] + dx * C.right[0] + dy * C.up[0], C.fwd[1] + dx * C.right[1] + dy * C.up[1], C.fwd[2] + dx * C.right[2] + dy * C.up[2]};
    if (inside) {
        const int n = s_n;
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            const int k = s_list[j];
            const RenderPrim& p = s_p[k];
            float nrm[3];
            const float t = p.kind == 0 ? hit_box(p, s_eye[k], d, nrm) : hit_cyl(p, s_eye[k], d, nrm);
            if (t > 0.f && t < best) { best = t; id = p.id; bn[0] = nrm[0]; bn[1] = nrm[1]; bn[2] = nrm[2]; }
        }
    }
    const bool hit = best >= C.near && best <= C.far;
    // OpenGL depth buffer value of the eye depth (what getCameraImage returns): z_ndc = ((f + n) - 2 f n / z) / (f - n), d = (z_ndc + 1) / 2
    const float zn = hit ? ((C.far + C.near) - 2.0f * C.far * C.near / best) / (C.far - C.near) : 1.0f;
    const float db = hit ? 0.5f * (zn + 1.0f) : 1.0f;
    const size_t g = (size_t)env * npix + (size_t)row * C.width + col;
    if (inside) {
        if (depth) depth[g] = db;
        if (seg) seg[g] = hit ? (unsigned char)id : 0;
        if (rgba) {
            uchar4 c = make_uchar4(C.background[0], C.background[1], C.background[2], 255);
            if (hit) {
                const unsigned char* base = C.color[id < RENDER_ID_ROBOT ? id : RENDER_ID_ROBOT];
                const float sh = 0.45f + 0.55f * fmaxf(0.f, bn[0] * C.light[0] + bn[1] * C.light[1] + bn[2] * C.light[2]);
                c = make_uchar4((unsigned char)(base[0] * sh), (unsigned char)(base[1] * sh), (unsigned char)(base[2] * sh), 255);
            }
            rgba[g] = c;
        }
    }
    if (points) {
        // the reference's deprojection (pybullet.py:213-241): NDC of the pixel CORNER (np.mgrid[-1:1:2/h, -1:1:2/w], y flipped) and of the
        // depth buffer, through inv(P V); then "infinite depth" (buffer >= 0.99) and the workspace box 0 < z < 0.67, -0.5 < x < 0.2
        const float xc = col * (2.0f / C.width) - 1.0f, yc = -(row * (2.0f / C.height) - 1.0f);
        const float ze = best;                                        // eye depth of this pixel
        const float ex = xc * C.tan_half_fov * C.aspect * ze, ey = yc * C.tan_half_fov * ze;
        const float px = C.eye[0] + ze * C.fwd[0] + ex * C.right[0] + ey * C.up[0], py = C.eye[1] + ze * C.fwd[1] + ex * C.right[1] + ey * C.up[1],
                    pz = C.eye[2] + ze * C.fwd[2] + ex * C.right[2] + ey * C.up[2];
        bool ok = hit && db < 0.99f;
        if (C.crop) ok = ok && pz > 0.0f && pz < 0.67f && px > -0.5f && px < 0.2f;
        const float qnan = __int_as_float(0x7fc00000);
        s_pts[3 * tid] = ok ? px : qnan; s_pts[3 * tid + 1] = ok ? py : qnan; s_pts[3 * tid + 2] = ok ? pz : qnan;
        if (valid && inside) valid[g] = ok;
        __syncthreads();
        // a tile row is RT pixels = 3 RT consecutive floats of the [N, H, W, 3] array
        const int wcols = min(RT, C.width - x0);
        for (int i = tid; i < RT * RT * 3; i += RT * RT) {
            const int r = i / (3 * RT), o = i - r * 3 * RT;
            if (y0 + r < C.height && o < 3 * wcols) points[((size_t)env * npix + (size_t)(y0 + r) * C.width + x0) * 3 + o] = s_pts[i];
        }
    }
}

template <typename T> void launch_render_setup(const EnvDev<T>& E, const RenderScene& R, RenderPrim* prims, cudaStream_t st) {
    render_setup_kernel<T><<<(E.n + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(E, R, prims);
    g_launches++;
}
void launch_render(const RenderPrim* prims, const RenderCamera& C, int n_envs, float* depth, unsigned char* rgba, unsigned char* seg, float* points, unsigned char* valid, cudaStream_t st) {
    const size_t npix = (size_t)C.width * C.height;
    for (int e0 = 0; e0 < n_envs; e0 += 65535) {        // gridDim.z <= 65535
        const int ne = n_envs - e0 < 65535 ? n_envs - e0 : 65535;
        const dim3 grid((C.width + RT - 1) / RT, (C.height + RT - 1) / RT, ne);
        render_kernel<<<grid, RT * RT, 0, st>>>(prims + (size_t)e0 * RENDER_MAX_PRIMS, C, depth ? depth + e0 * npix : nullptr,
                                                 rgba ? reinterpret_cast<uchar4*>(rgba) + e0 * npix : nullptr, seg ? seg + e0 * npix : nullptr,
                                                 points ? points + 3 * e0 * npix : nullptr, valid ? valid + e0 * npix : nullptr);
        g_launches++;
    }
}
template void launch_render_setup<float>(const EnvDev<float>&, const RenderScene&, RenderPrim*, cudaStream_t);
template void launch_render_setup<double>(const EnvDev<double>&, const RenderScene&, RenderPrim*, cudaStream_t);

}  // namespace pg
