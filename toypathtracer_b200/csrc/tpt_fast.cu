// Fast (throughput) mode kernels: same estimator as the reference's Trace/Scatter (Cpp/Source/Test.cpp:83-234),
// one independent XorShift32 stream per (pixel, frame) instead of the reference's per-row stream, FMA
// contraction and fast intrinsics allowed. Results agree with the reference statistically (same expectation,
// same rays/sample), not bitwise — bitwise parity is the exact mode's job (tpt_exact.cu).
//
// variant 0  "megakernel": one thread per pixel, loops over frames x spp (what the reference's own GPU shaders
//            do, Cpp/Windows/ComputeShader.hlsl:353-395). Baseline for the persistent design.
// variant 1  "persistent wavefront": one CTA per SM slot, resident for the whole draw. The CTA pulls tiles of
//            TILE_PIX pixels from a global counter; inside a tile every lane runs a path state machine and
//            refills itself with the next (pixel, sample) from a warp-aggregated shared counter as soon as its
//            path ends (ray regeneration), so the ray-vs-all-spheres sweep — 85 % of the instructions — always
//            runs with full warps. Radiance is accumulated per pixel in shared memory; finished tiles are
//            written with coalesced 128-bit stores (and 128-bit loads of `prev` when accumulating).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tpt_integrator.cuh"
#include "tpt_device_utils.cuh"
#include "tpt_launch.h"
#include "tpt_fastdiv.h"

namespace tpt {

constexpr int kFastThreads = 256;
constexpr int kTilePix = 1024;          // pixels per tile (12 KB of float3 accumulators)
constexpr int kMaxFramesPerDraw = 256;  // per-frame blend weights live in shared memory

__device__ __forceinline__ uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t pixel_seed(uint32_t pixelIndex, uint32_t frame)
{
    return fmix32(pixelIndex * 0x9E3779B9u + fmix32(frame + 0x7F4A7C15u)) | 1u;
}

// Per-frame weights of the progressive blend (Test.cpp:272-276,293-295) when `numFrames` frames are fused in
// one launch: image = prev*wPrev + sum_f mean_f * w[f], w[f] = (1-lerp_f) * prod_{g>f} lerp_g.
// Fast mode draws its lens and roughness samples analytically instead of by rejection (Maths.cpp:20-37): the same
// distributions (uniform on the unit disk / in the unit ball) from a fixed number of draws, so no lane of a warp waits in
// a rejection loop for its neighbours (ncu: the two loops were 23 + ~25 warp instructions per path step at 18 / 3.5 active
// lanes). The reference's own GPU ports sample analytically as well (ComputeShader.hlsl:18-35). TPT_FAST_ANALYTIC=0
// restores the rejection loops for A/B runs.
#ifndef TPT_FAST_ANALYTIC
#define TPT_FAST_ANALYTIC 1
#endif
__device__ __forceinline__ int opaque_int(int x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ V3 fast_in_unit_disk(uint32_t& state)
{
#if TPT_FAST_ANALYTIC
    const float r = M<false>::sqrt_(RandomFloat01(state));
    float sa, ca;
    __sincosf(2.0f * TPT_PI * RandomFloat01(state), &sa, &ca);
    return v3(r * ca, r * sa, 0.0f);
#else
    return RandomInUnitDisk(state);
#endif
}
__device__ __forceinline__ V3 fast_in_unit_sphere(uint32_t& state)
{
#if TPT_FAST_ANALYTIC
    const float z = 1.0f - 2.0f * RandomFloat01(state);
    float sa, ca;
    __sincosf(2.0f * TPT_PI * RandomFloat01(state), &sa, &ca);
    const float rad = __powf(RandomFloat01(state), 1.0f / 3.0f);                 // radius = u^(1/3) as ex2(lg2(u)/3); u = 0 -> 0
    const float r = M<false>::sqrt_(fmaxf(1.0f - z * z, 0.0f)) * rad;
    return v3(r * ca, r * sa, z * rad);
#else
    return RandomInUnitSphere(state);
#endif
}
// Maths.h:437-442 with the analytic lens sample
__device__ __forceinline__ Ray fast_get_ray(const Camera88& c, float s, float t, uint32_t& state)
{
    V3 rd = c.lensRadius * fast_in_unit_disk(state);
    V3 offset = ld3(c.uu) * rd.x + ld3(c.vv) * rd.y;
    Ray r;
    r.orig = ld3(c.origin) + offset;
    r.dir = M<false>::normalize(ld3(c.lowerLeftCorner) + s * ld3(c.horizontal) + t * ld3(c.vertical) - ld3(c.origin) - offset);
    return r;
}

__device__ __forceinline__ void blend_weights(const DrawParams& p, float* w, float& wPrev)
{
    float suffix = 1.0f;
    for (int fi = p.numFrames - 1; fi >= 0; --fi)
    {
        float lf = lerp_fac(p.frame0 + fi, p.flags);
        w[fi] = (1.0f - lf) * suffix;
        suffix *= lf;
    }
    wPrev = suffix;
}

// prev*wPrev + acc for one pixel (Test.cpp:293-295 with the frames of this draw folded into acc). A zero weight must
// not read `prev` as a number (NaN * 0), but the pixel's alpha is still the buffer's: the reference never writes it
// (Maths.h:38).
__device__ __forceinline__ float4 blend_prev(const DrawParams& p, const float* px, float wPrev, float ax, float ay, float az)
{
    if (wPrev != 0.0f)
    {
        float4 prev = ld_stream_f4(px);
        prev.x = prev.x * wPrev + ax; prev.y = prev.y * wPrev + ay; prev.z = prev.z * wPrev + az;
        return prev;
    }
    return make_float4(ax, ay, az, p.zeroAlpha ? 0.0f : __ldg(px + 3));
}

// ---- variant 0 ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFastThreads)
k_fast_mega(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights, uint32_t stagedBytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ float sW[kMaxFramesPerDraw];
    __shared__ float sWPrev;
    stage_blob(smem, blob, stagedBytes, &bar);
    if (threadIdx.x == 0) { float wp; blend_weights(p, sW, wp); sWPrev = wp; }
    __syncthreads();
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);

    // 16x16 pixel block, each warp an 8x4 patch
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tilesX = (p.width + 15) / 16;
    const int tx = blockIdx.x % tilesX, ty = blockIdx.x / tilesX;
    const int x = tx * 16 + (warp & 1) * 8 + (lane & 7);
    const int ri = ty * 16 + (warp >> 1) * 4 + (lane >> 3);
    unsigned rc = 0;
    if (x < p.width && ri < p.numRows)
    {
        const int y = p.row0 + ri * p.rowStep;
        SerialHitter<false> hitter;
        V3 acc = v3(0, 0, 0);
        const float invSpp = 1.0f / (float)p.spp;
        for (int fi = 0; fi < p.numFrames; ++fi)
        {
            uint32_t state = pixel_seed((uint32_t)(y * p.width + x), (uint32_t)(p.frame0 + fi));
            V3 col = v3(0, 0, 0);
            for (int s = 0; s < p.spp; ++s)
            {
                float u = ((float)x + RandomFloat01(state)) * p.invWidth;
                float v = ((float)y + RandomFloat01(state)) * p.invHeight;
                Ray r = fast_get_ray(p.cam, u, v, state);
                col = col + trace_fast(sc, r, state, rc, hitter);
            }
            acc = acc + col * (invSpp * sW[fi]);
        }
        float* px = p.image + ((size_t)(p.packed ? ri : y) * p.width + x) * 4;
        st_stream_f4(px, blend_prev(p, px, sWPrev, acc.x, acc.y, acc.z));
    }
    // one atomic per warp
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(p.rayCounter, (unsigned long long)rc);
}

// ---- variant 1 ------------------------------------------------------------------------------------------------
struct PathState
{
    V3 o, d;          // ray to intersect next
    V3 thr, col;      // throughput before the current vertex, radiance so far
    V3 nextDir;       // Lambert bounce direction, pending while shadow rays are in flight
    V3 thrAlb;        // thr * albedo at the pending Lambert vertex
    V3 nl;            // shading normal facing the incoming ray (Test.cpp:129)
    V3 pend;          // radiance the in-flight shadow ray carries if it reaches its light
    V3 albedo;
    uint32_t rng;
    int pix;          // pixel index inside the tile
    float weight;
    int kind;         // 0: path ray, 1+j: shadow ray towards light j
    int depth;
    int mid;          // sphere id of the pending Lambert vertex (skip-self test, Test.cpp:100)
    bool doMaterialE;
    bool active;
};

template <int MINB>
__global__ void __launch_bounds__(kFastThreads, MINB)
k_fast_persistent(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights,
                  uint32_t stagedBytes, int numTiles)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ float sW[kMaxFramesPerDraw];
    __shared__ float sWPrev;
    __shared__ float sAcc[kTilePix * 3];
    __shared__ int sTile;
    __shared__ int sTask;
    stage_blob(smem, blob, stagedBytes, &bar);
    if (threadIdx.x == 0) { float wp; blend_weights(p, sW, wp); sWPrev = wp; }
    for (int i = threadIdx.x; i < kTilePix * 3; i += kFastThreads) sAcc[i] = 0.0f;
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);
    const int lane = threadIdx.x & 31;
    const long long regionPix = (long long)p.numRows * p.width;
    const int S = p.spp * p.numFrames;          // samples per pixel in this draw
    const float invSpp = 1.0f / (float)p.spp;
    SerialHitter<false> hitter;
    unsigned rc = 0;

    for (;;)
    {
        __syncthreads(); // sAcc zeroed / previous tile written
        if (threadIdx.x == 0) { sTile = (int)atomicAdd(p.workCounter, 1u); sTask = 0; }
        __syncthreads();
        const int tile = sTile;
        if (tile >= numTiles) break;
        const long long pix0 = (long long)tile * kTilePix;
        const int tilePix = (int)(regionPix - pix0 < kTilePix ? regionPix - pix0 : kTilePix);
        const int tileTasks = tilePix * S;

        PathState st;
        st.active = false;
        for (;;)
        {
            // ---- regeneration: lanes without a path take the next (pixel, sample) of the tile
            unsigned need = __ballot_sync(0xffffffffu, !st.active);
            if (need)
            {
                int base = 0;
                const int leader = __ffs(need) - 1;
                if (lane == leader) base = atomicAdd(&sTask, __popc(need));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!st.active)
                {
                    const int task = base + __popc(need & ((1u << lane) - 1u));
                    if (task < tileTasks)
                    {
                        // samples of one pixel sit in adjacent lanes: coherent primary rays
                        const int pixIn = task / S, sIdx = task % S;
                        const int fi = sIdx / p.spp;
                        const long long gp = pix0 + pixIn;
                        const int ri = (int)(gp / p.width), x = (int)(gp % p.width);
                        const int y = p.row0 + ri * p.rowStep;
                        st.rng = pixel_seed((uint32_t)(y * p.width + x) * (uint32_t)p.spp + (uint32_t)(sIdx % p.spp),
                                            (uint32_t)(p.frame0 + fi));
                        float u = ((float)x + RandomFloat01(st.rng)) * p.invWidth;
                        float v = ((float)y + RandomFloat01(st.rng)) * p.invHeight;
                        Ray r = fast_get_ray(p.cam, u, v, st.rng);
                        st.o = r.orig; st.d = r.dir;
                        st.thr = v3(1, 1, 1); st.col = v3(0, 0, 0);
                        st.pix = pixIn; st.weight = invSpp * sW[fi];
                        st.kind = 0; st.depth = 0; st.doMaterialE = true; st.active = true;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, st.active)) break;

            // ---- intersect: every active lane sweeps all spheres (shared-memory broadcast reads)
            float t = TPT_MAX_T;
            int id = -1;
            if (st.active) { id = hitter.hit(sc, st.o, st.d, TPT_MIN_T, TPT_MAX_T, t); ++rc; }

            // ---- shade
            bool wantLight = false;   // lane must pick its next shadow ray (or resume the path)
            int lightFrom = 0;
            bool finished = false;
            if (st.active)
            {
                if (st.kind == 0)
                {
                    if (id < 0) { st.col = st.col + st.thr * sky(st.d, sc); finished = true; }
                    else
                    {
                        Q4 s = ld_sph(sc, id);
                        V3 pos = st.o + st.d * t;
                        V3 normal = (pos - v3(s.x, s.y, s.z)) * sc.invRadius[id];
                        const int mid = id < sc.count ? id : sc.count;
                        Mat mat = load_mat(sc, mid);
                        if (st.depth >= TPT_MAX_DEPTH) { st.col = st.col + st.thr * mat.emissive; finished = true; }
                        else if (mat.type == kLambert)
                        {
                            if (st.doMaterialE) st.col = st.col + st.thr * mat.emissive;
                            V3 target = normal + RandomUnitVector<false>(st.rng);
                            st.nextDir = M<false>::normalize(target);
                            st.thrAlb = st.thr * mat.albedo;
                            st.albedo = mat.albedo;
                            st.nl = dot(normal, st.d) < 0.0f ? normal : neg(normal);
                            st.mid = mid;
                            st.o = pos;
                            wantLight = true; lightFrom = 0;
                        }
                        else
                        {
                            V3 att, outDir;
                            bool ok = scatter_specular<false>(mat, st.d, pos, normal, st.rng, att, outDir);
                            if (!ok) { st.col = st.col + st.thr * mat.emissive; finished = true; }
                            else
                            {
                                if (st.doMaterialE) st.col = st.col + st.thr * mat.emissive;
                                st.doMaterialE = true;
                                st.thr = st.thr * att;
                                st.o = pos; st.d = outDir; ++st.depth;
                            }
                        }
                    }
                }
                else
                {
                    const int j = st.kind - 1;
                    if (id == sc.lights[j].id) st.col = st.col + st.pend;
                    wantLight = true; lightFrom = j + 1;
                }
            }
            // Lambert vertices and returning shadow rays converge here: next light sample or resume the path
            if (wantLight)
            {
                int j = lightFrom;
                while (j < sc.nLights && sc.lights[j].id == st.mid) ++j;
                if (j < sc.nLights)
                {
                    const LightRec Lr = sc.lights[j];
                    // sample_light (Test.cpp:104-131) with the facing normal already resolved
                    V3 scn = v3(Lr.cx, Lr.cy, Lr.cz);
                    V3 sw = M<false>::normalize(scn - st.o);
                    V3 su = M<false>::normalize(cross(fabsf(sw.x) > 0.01f ? v3(0, 1, 0) : v3(1, 0, 0), sw));
                    V3 sv = cross(sw, su);
                    V3 pc = st.o - scn;
                    float cosAMax = M<false>::sqrt_(1.0f - __fdividef(Lr.radius * Lr.radius, dot(pc, pc)));
                    float eps1 = RandomFloat01(st.rng), eps2 = RandomFloat01(st.rng);
                    float cosA = 1.0f - eps1 + eps1 * cosAMax;
                    float sinA = M<false>::sqrt_(1.0f - cosA * cosA);
                    float phi = 2.0f * TPT_PI * eps2;
                    float sp, cp;
                    __sincosf(phi, &sp, &cp);
                    V3 l = su * (cp * sinA) + sv * (sp * sinA) + sw * cosA;
                    float omega = 2.0f * TPT_PI * (1.0f - cosAMax);
                    float dl = dot(l, st.nl);
                    float m = (0.0f < dl) ? dl : 0.0f;
                    st.pend = st.thr * ((st.albedo * v3(Lr.ex, Lr.ey, Lr.ez)) * (m * omega * (1.0f / TPT_PI)));
                    st.d = l;
                    st.kind = 1 + j;
                }
                else
                {
                    st.d = st.nextDir;
                    st.thr = st.thrAlb;
                    st.kind = 0;
                    st.doMaterialE = false;
                    ++st.depth;
                }
            }
            if (finished)
            {
                atomicAdd(&sAcc[st.pix * 3 + 0], st.col.x * st.weight);
                atomicAdd(&sAcc[st.pix * 3 + 1], st.col.y * st.weight);
                atomicAdd(&sAcc[st.pix * 3 + 2], st.col.z * st.weight);
                st.active = false;
            }
        }

        // ---- tile done: coalesced 128-bit write-out (+ 128-bit read of prev when accumulating)
        __syncthreads();
        const float wPrev = sWPrev;
        for (int i = threadIdx.x; i < tilePix; i += kFastThreads)
        {
            const long long gp = pix0 + i;
            const int ri = (int)(gp / p.width), x = (int)(gp % p.width);
            const int y = p.row0 + ri * p.rowStep;
            float* px = p.image + ((size_t)(p.packed ? ri : y) * p.width + x) * 4;
            st_stream_f4(px, blend_prev(p, px, wPrev, sAcc[i * 3 + 0], sAcc[i * 3 + 1], sAcc[i * 3 + 2]));
            sAcc[i * 3 + 0] = 0.0f; sAcc[i * 3 + 1] = 0.0f; sAcc[i * 3 + 2] = 0.0f;
        }
    }
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(p.rayCounter, (unsigned long long)rc);
}


// ---- variant 3 ------------------------------------------------------------------------------------------------
// "persistent queue": no tiles, no block barriers after the prologue. Every WARP pulls slabs of kSlabPix pixels x
// one sample index from a global counter (one atomic per 128 paths), deals the slab's paths to its idle lanes
// (ray regeneration, no atomics inside the warp: the slab cursor is warp-uniform), and every finished path adds
// its weighted radiance to the float4 accumulation buffer with ONE 128-bit vector reduction
// (red.global.add.v4.f32, SASS REDG.E.ADD.F32x4) that resolves in L2. A tiny prepare kernel first scales the
// buffer by the weight of `prev` (or zeroes it). End-of-kernel tail = one slab per warp instead of one 1024-pixel
// tile per CTA, which matters at 1280x720x4spp where a warp's share of the whole frame is only ~1000 paths.
#ifndef TPT_SLAB_PIX
#define TPT_SLAB_PIX 64
#endif
constexpr int kSlabPix = TPT_SLAB_PIX;     // paths per slab (pixels x one sample index); measured at 1280x720x4spp: 32 -> 20.1, 64 -> 20.3, 128 -> 19.5, 256 -> 18.0 Gray/s
#ifndef TPT_QUEUE_THREADS
#define TPT_QUEUE_THREADS 128
#define TPT_QUEUE_MINB 6
#endif
#ifndef TPT_BIG_THREADS
#define TPT_BIG_THREADS 768      // one CTA per SM for scenes that fill the shared memory (A/B: 1024 = 32 warps at 64 registers)
#endif
#ifndef TPT_P1_GROUP
#define TPT_P1_GROUP 16
#endif
constexpr int kQueueThreads = TPT_QUEUE_THREADS;   // measured at 1280x720x4spp: 64 -> 20.33, 128 -> 20.33, 256 -> 20.46 Gray/s

__global__ void k_prepare_image(DrawParams p, float wPrev)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)p.numRows * p.width;
    if (idx >= n) return;
    const int ri = (int)(idx / p.width), x = (int)(idx % p.width);
    const int y = p.row0 + ri * p.rowStep;
    float* px = p.image + ((size_t)(p.packed ? ri : y) * p.width + x) * 4;
    float4 v = blend_prev(p, px, wPrev, 0.0f, 0.0f, 0.0f);
    *reinterpret_cast<float4*>(px) = v;   // stays in L2 for the reductions that follow
}

__device__ __forceinline__ void red_add_f4(float* addr, float x, float y, float z)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(0.0f) : "memory");
}

// Camera rays of one slab (<= kSlabPix consecutive pixels of the region, one sample index), generated by the whole warp
// in one convergent burst (Test.cpp:286-288 + Maths.h:437-442) into the warp's shared buffer:
// {origin.xyz, rng state} {direction.xyz, tag}; tag = image float4 offset (tileBase < 0) or tileBase + q.
__device__ __forceinline__ void generate_slab_rays(const DrawParams& p, float4 (*rays)[2], int lane, uint32_t count,
                                                   int x0, int ri0, uint32_t sample, uint32_t frame, int tileBase)
{
    for (uint32_t q = (uint32_t)lane; q < count; q += 32)
    {
        int x = x0 + (int)q, ri = ri0;
        while (x >= p.width) { x -= p.width; ++ri; }
        const int y = p.row0 + ri * p.rowStep;
        uint32_t rng = pixel_seed((uint32_t)(y * p.width + x) * (uint32_t)p.spp + sample, frame);
        float u = ((float)x + RandomFloat01(rng)) * p.invWidth;
        float v = ((float)y + RandomFloat01(rng)) * p.invHeight;
        Ray r = fast_get_ray(p.cam, u, v, rng);
        const uint32_t tag = tileBase < 0 ? (uint32_t)((p.packed ? ri : y) * p.width + x) : (uint32_t)tileBase + q;
        rays[q][0] = make_float4(r.orig.x, r.orig.y, r.orig.z, __uint_as_float(rng));
        rays[q][1] = make_float4(r.dir.x, r.dir.y, r.dir.z, __uint_as_float(tag));
    }
}

struct QPath
{
    V3 o, d;
    V3 thr, col;
    V3 nextDir;
    V3 thrAlb;
    V3 nl;
    V3 pend;
    V3 albedo;
    uint32_t rng;
    uint32_t pixOff;  // float4 index of the pixel in the image buffer
    float weight;
    int kind;
    int depth;
    int mid;
    bool doMaterialE;
    bool active;
};

// Fast-mode sweep in expanded form: with od = o.d, oo = o.o and K = |s|^2 - r^2 (per sphere, computed once per CTA in
// double precision) nb = s.d - od and c = K + oo - 2 s.o, i.e. 8 instead of 10 FP32 issue slots per (ray, sphere):
//   nb = fma(sx,dx, fma(sy,dy, fma(sz,dz, -od)));  c = fma(sx,-2ox, fma(sy,-2oy, fma(sz,-2oz, K+oo)));  discr = fma(nb,nb,-c)
// Same two-pass structure and the same behind-the-origin rejection as SerialHitter. Algebraically the reference's
// test (Maths.cpp:97-102); rounding differs (c loses ~|s|^2 * 2^-24 absolute accuracy), which the fast mode's
// statistical tests bound. Padded "impossible" spheres get K = +1e30 so they can never be candidates here.
struct FastHitterK
{
    uint32_t sphK;      // shared-memory address of {sx, sy, sz, K}[simdCount]
    int simdCount;
    __device__ __forceinline__ float4 ld(int i) const
    {
        float4 r;
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sphK + (uint32_t)i * 16u));
        return r;
    }
    __device__ __forceinline__ int hit(const SceneView&, V3 o, V3 d, float tMin, float tMax, float& tOut) const
    {
        const float nod = -fmaf(o.x, d.x, fmaf(o.y, d.y, o.z * d.z));
        const float oo = fmaf(o.x, o.x, fmaf(o.y, o.y, o.z * o.z));
        const float ax = -2.0f * o.x, ay = -2.0f * o.y, az = -2.0f * o.z;
        float bestT = tMax;
        int bestId = -1;
        for (int base = 0; base < simdCount; base += 32)
        {
            const int n = simdCount - base < 32 ? simdCount - base : 32;
            uint32_t neg = 0;
#pragma unroll
            for (int k = 0; k < 32; k += 4)
            {
                if (k < n)
                {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        const float4 s = ld(base + k + j);
                        const float nb = fmaf(s.x, d.x, fmaf(s.y, d.y, fmaf(s.z, d.z, nod)));
                        const float c = fmaf(s.x, ax, fmaf(s.y, ay, fmaf(s.z, az, s.w + oo)));
                        const float discr = fmaf(nb, nb, -c);
                        // reject: discr < 0, or centre behind (nb < 0) with the origin outside (c > 0)
                        const uint32_t rej = __float_as_uint(discr) | (__float_as_uint(nb) & ~__float_as_uint(c));
                        neg = __funnelshift_l(rej, neg, 1);
                    }
                }
            }
            uint32_t cand = ~neg & (n == 32 ? 0xffffffffu : ((1u << n) - 1u));
            while (cand)
            {
                const int bit = 31 - __clz((int)cand);
                cand &= ~(1u << bit);
                const int i = base + (n - 1 - bit);
                const float4 s = ld(i);
                const float nb = fmaf(s.x, d.x, fmaf(s.y, d.y, fmaf(s.z, d.z, nod)));
                const float c = fmaf(s.x, ax, fmaf(s.y, ay, fmaf(s.z, az, s.w + oo)));
                const float discr = fmaf(nb, nb, -c);
                if (discr > 0.0f)
                {
                    const float sq = M<false>::sqrt_(discr);
                    float t = nb - sq;
                    if (t <= tMin) t = nb + sq;
                    if (t > tMin && t < bestT) { bestT = t; bestId = i; }
                }
            }
        }
        tOut = bestT;
        return bestId;
    }
};

// Rewrites the staged sphere array IN PLACE from {s, r^2} to {s, K} (one CTA-wide pass, double precision for the
// cancellation |s|^2 - r^2). The fast kernels need r^2 nowhere else (the normal uses the centre and invRadius), so the
// expanded form costs no extra shared memory and works for any sphere count.
__device__ __forceinline__ void build_sphK(const SceneView& sc, float4* sphInSmem)
{
    for (int i = threadIdx.x; i < sc.simdCount; i += blockDim.x)
    {
        const float4 s = sphInSmem[i];
        const double K = (double)s.x * s.x + (double)s.y * s.y + (double)s.z * s.z - (double)s.w;
        sphInSmem[i].w = i < sc.count ? (float)K : 1.0e30f;
    }
}

// Packed-pair form of the expanded sweep: Blackwell's fma.rn.f32x2 / add.rn.f32x2 (SASS FFMA2 / FADD2) evaluate TWO spheres
// per instruction, with the ray constants as scalar broadcast operands. The kernels are FP32-issue bound (ncu: issue slots
// 78 % busy, FMA pipe 42 %), so halving the issue slots of pass 1 is worth more than its FLOPs: per PAIR of spheres
//   2 LDS.128 {x0,x1,y0,y1} {z0,z1,-K0,-K1},  1 FADD2 (-K - o.o),  3 FFMA2 (nb),  3 FFMA2 (-c),  1 FFMA2 (discr),
//   2 LOP3 + 2 SHF (sign bits into the candidate mask)           = 7 issue slots per sphere instead of 12.5.
// -c and discr are the exact negations/equals of FastHitterK's c and discr (same products, IEEE negation symmetry), so pass
// 2 — unchanged, on the {s, K} array — sees the same candidates it would have computed itself.
__device__ __forceinline__ unsigned long long f2_bcast(float x) { unsigned long long r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x)); return r; }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// Packed-pair pass 1 shared by FastHitterK2 / FastHitterK2C: spheres [s0, s0+G) = G/2 pairs of the pair array at shared
// address sphP; shifts G rejection signs into `neg` (sphere s0 ends up at the highest of the G new bits).
struct SweepConsts { unsigned long long DX, DY, DZ, NOD, BX, BY, BZ, NOO; };
template <int G>
__device__ __forceinline__ void sweep_full(uint32_t sphP, int s0, uint32_t& neg, const SweepConsts& c)
{
#pragma unroll
    for (int j = 0; j < G; j += 2)
    {
        unsigned long long xx, yy, zz, kk;
        const uint32_t a = sphP + (uint32_t)((s0 + j) >> 1) * 32u;
        asm("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(xx), "=l"(yy) : "r"(a));
        asm("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(zz), "=l"(kk) : "r"(a + 16u));
        const unsigned long long nb = f2_fma(xx, c.DX, f2_fma(yy, c.DY, f2_fma(zz, c.DZ, c.NOD)));
        const unsigned long long negc = f2_fma(xx, c.BX, f2_fma(yy, c.BY, f2_fma(zz, c.BZ, f2_add(kk, c.NOO))));
        const unsigned long long discr = f2_fma(nb, nb, negc);
        // reject: discr < 0, or centre behind (nb < 0) with the origin outside (-c < 0)
        const uint32_t r0 = (uint32_t)discr | ((uint32_t)nb & (uint32_t)negc);
        const uint32_t r1 = (uint32_t)(discr >> 32) | ((uint32_t)(nb >> 32) & (uint32_t)(negc >> 32));
        neg = __funnelshift_l(r0, neg, 1);
        neg = __funnelshift_l(r1, neg, 1);
    }
}
// The straight-line group between two (warp-uniform) bounds checks is the instruction scheduler's window: a group of 16
// spheres keeps eight independent packed chains in flight (measured at 1280x720x4spp: groups of 4 / 8 / 16 spheres ->
// 21.6 / 21.9 / 22.2 Gray/s). Counts that are not a multiple of the group fall through to halved groups, down to 4.
template <int G>
__device__ __forceinline__ void sweep_upto(uint32_t sphP, int s0, int left, uint32_t& neg, const SweepConsts& c)
{
    if (left >= G) sweep_full<G>(sphP, s0, neg, c);
    else if constexpr (G > 4)
    {
        if (left > 0)
        {
            sweep_upto<G / 2>(sphP, s0, left, neg, c);
            sweep_upto<G / 2>(sphP, s0 + G / 2, left - G / 2, neg, c);
        }
    }
}
// 64 spheres [base, base+n) -> candidate mask, sphere base + J <-> bit 63 - J (n a multiple of 4, 4 <= n <= 64)
__device__ __forceinline__ unsigned long long sweep_chunk64(uint32_t sphP, int base, int n, const SweepConsts& c)
{
    const int n0 = n < 32 ? n : 32, n1 = n - n0;
    uint32_t neg0 = 0, neg1 = 0;
#pragma unroll
    for (int k = 0; k < 32; k += TPT_P1_GROUP) sweep_upto<TPT_P1_GROUP>(sphP, base + k, n0 - k, neg0, c);
    if (n1 > 0)
    {
#pragma unroll
        for (int k = 0; k < 32; k += TPT_P1_GROUP) sweep_upto<TPT_P1_GROUP>(sphP, base + 32 + k, n1 - k, neg1, c);
    }
    // left-align: after n shifts sphere j of the round sits at bit n-1-j
    const uint32_t c0 = ~neg0 << (32 - n0);
    const uint32_t c1 = n1 > 0 ? ~neg1 << (32 - n1) : 0u;
    return ((unsigned long long)c0 << 32) | c1;
}
struct FastHitterK2
{
    uint32_t sphK;      // shared-memory address of {sx, sy, sz, K}[simdCount]         (pass 2)
    uint32_t sphP;      // shared-memory address of the pair array, 32 B per pair       (pass 1)
    int simdCount;
    __device__ __forceinline__ float4 ld(int i) const
    {
        float4 r;
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sphK + (uint32_t)i * 16u));
        return r;
    }
    __device__ __forceinline__ void ldp(int pair, unsigned long long& xx, unsigned long long& yy, unsigned long long& zz, unsigned long long& kk) const
    {
        asm("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(xx), "=l"(yy) : "r"(sphP + (uint32_t)pair * 32u));
        asm("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(zz), "=l"(kk) : "r"(sphP + (uint32_t)pair * 32u + 16u));
    }
    __device__ __forceinline__ int hit(const SceneView&, V3 o, V3 d, float tMin, float tMax, float& tOut) const
    {
        const float nod = -fmaf(o.x, d.x, fmaf(o.y, d.y, o.z * d.z));
        const float oo = fmaf(o.x, o.x, fmaf(o.y, o.y, o.z * o.z));
        const float ax = -2.0f * o.x, ay = -2.0f * o.y, az = -2.0f * o.z;
        SweepConsts C;
        C.DX = f2_bcast(d.x); C.DY = f2_bcast(d.y); C.DZ = f2_bcast(d.z); C.NOD = f2_bcast(nod);
        C.BX = f2_bcast(-ax); C.BY = f2_bcast(-ay); C.BZ = f2_bcast(-az); C.NOO = f2_bcast(-oo);
        float bestT = tMax;
        int bestId = -1;
        // chunks of 64 spheres: pass 1 fills a 64-bit candidate mask (sphere base + J <-> bit 63 - J), pass 2 walks it
        for (int base = 0; base < simdCount; base += 64)
        {
            unsigned long long cand = sweep_chunk64(sphP, base, simdCount - base < 64 ? simdCount - base : 64, C);
            // pass 2, two candidates per trip (independent LDS.128 / 9 FMA / MUFU chains); ascending sphere order as before
            while (cand)
            {
                const int ja = __clzll((long long)cand);
                cand &= ~(0x8000000000000000ull >> ja);
                const bool hasB = cand != 0ull;
                const int jb = hasB ? __clzll((long long)cand) : ja;
                cand &= ~(0x8000000000000000ull >> jb);
                const float4 sa = ld(base + ja), sb = ld(base + jb);
                const float nba = fmaf(sa.x, d.x, fmaf(sa.y, d.y, fmaf(sa.z, d.z, nod)));
                const float nbb = fmaf(sb.x, d.x, fmaf(sb.y, d.y, fmaf(sb.z, d.z, nod)));
                const float ca = fmaf(sa.x, ax, fmaf(sa.y, ay, fmaf(sa.z, az, sa.w + oo)));
                const float cb = fmaf(sb.x, ax, fmaf(sb.y, ay, fmaf(sb.z, az, sb.w + oo)));
                const float da = fmaf(nba, nba, -ca), db = fmaf(nbb, nbb, -cb);
                const float sqa = M<false>::sqrt_(fmaxf(da, 0.0f)), sqb = M<false>::sqrt_(fmaxf(db, 0.0f));
                float ta = nba - sqa, tb = nbb - sqb;
                if (ta <= tMin) ta = nba + sqa;
                if (tb <= tMin) tb = nbb + sqb;
                if (da > 0.0f && ta > tMin && ta < bestT) { bestT = ta; bestId = base + ja; }
                if (hasB && db > 0.0f && tb > tMin && tb < bestT) { bestT = tb; bestId = base + jb; }
            }
        }
        tOut = bestT;
        return bestId;
    }
};

// Conservative form for scenes the expanded form is NOT accurate enough for (centres far from the origin: the gate in
// tpt_set_scene fails, e.g. the 4096-sphere stress scene): pass 1 is the packed expanded-form sweep with an error bound
// folded into its additive constants, so it can only ADD candidates; pass 2 evaluates every candidate in the reference
// form (Maths.cpp:97-102) on the untouched {s, r^2} array, so hit decisions and distances are the reference form's
// (tpt_debug_hit: bit-identical ids and distances on millions of rays). Bound (u = 2^-24, |d| = 1): |d nb| <= 6u(|s|+|o|),
// |d c| <= 9u(|s|+|o|)^2, hence |d discr| <= 22u(|s|+|o|)^2 for the expanded form and <= 12u(|s|+|o|)^2 for the reference
// form; with (|s|+|o|)^2 <= 2(|s|^2 + o.o) the margin eps_i = 2^-17 (|s_i|^2 + o.o) = 64u * 2(...) covers both. It splits
// into a per-sphere constant (stored with -K_i) and a per-ray constant (folded into -o.o): no extra instruction, and a
// far-away sphere (the ground, |s| = 1000) does not loosen the test of the others. The "wholly behind" rejection needs no
// margin of its own: a sphere it drops wrongly has a true nb below 6u(|s|+|o|) << tMin/2, i.e. both roots below tMin.
struct FastHitterK2C
{
    uint32_t sph;       // shared-memory address of the ORIGINAL {sx, sy, sz, r^2}[simdCount]   (pass 2)
    uint32_t sphP;      // pair array {x0,x1,y0,y1}{z0,z1,-K0,-K1}                             (pass 1)
    int simdCount;
    __device__ __forceinline__ int hit(const SceneView&, V3 o, V3 d, float tMin, float tMax, float& tOut) const
    {
        const float nod = -fmaf(o.x, d.x, fmaf(o.y, d.y, o.z * d.z));
        const float oo = fmaf(o.x, o.x, fmaf(o.y, o.y, o.z * o.z));
        SweepConsts C;
        C.DX = f2_bcast(d.x); C.DY = f2_bcast(d.y); C.DZ = f2_bcast(d.z); C.NOD = f2_bcast(nod);
        C.BX = f2_bcast(2.0f * o.x); C.BY = f2_bcast(2.0f * o.y); C.BZ = f2_bcast(2.0f * o.z);
        C.NOO = f2_bcast(oo * 7.62939453125e-6f - oo);     // the ray's share of the margin: 2^-17 o.o
        float bestT = tMax;
        int bestId = -1;
        for (int base = 0; base < simdCount; base += 64)
        {
            unsigned long long cand = sweep_chunk64(sphP, base, simdCount - base < 64 ? simdCount - base : 64, C);
            // pass 2 on the untouched {s, r^2} array in the reference form (Maths.cpp:97-102), two candidates per trip
            while (cand)
            {
                const int ja = __clzll((long long)cand);
                cand &= ~(0x8000000000000000ull >> ja);
                const bool hasB = cand != 0ull;
                const int jb = hasB ? __clzll((long long)cand) : ja;
                cand &= ~(0x8000000000000000ull >> jb);
                Q4 sa, sb;
                asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sa.x), "=f"(sa.y), "=f"(sa.z), "=f"(sa.w) : "r"(sph + (uint32_t)(base + ja) * 16u));
                asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sb.x), "=f"(sb.y), "=f"(sb.z), "=f"(sb.w) : "r"(sph + (uint32_t)(base + jb) * 16u));
                float nba, nbb;
                const float da = sphere_discr<false>(sa, o, d, nba), db = sphere_discr<false>(sb, o, d, nbb);
                const float sqa = M<false>::sqrt_(fmaxf(da, 0.0f)), sqb = M<false>::sqrt_(fmaxf(db, 0.0f));
                float ta = nba - sqa, tb = nbb - sqb;
                if (ta <= tMin) ta = nba + sqa;
                if (tb <= tMin) tb = nbb + sqb;
                if (da > 0.0f && ta > tMin && ta < bestT) { bestT = ta; bestId = base + ja; }
                if (hasB && db > 0.0f && tb > tMin && tb < bestT) { bestT = tb; bestId = base + jb; }
            }
        }
        tOut = bestT;
        return bestId;
    }
};

// Pair array for FastHitterK2C straight from the staged {s, r^2} array (which stays as it is): -K_i plus the sphere's share
// of the margin, 2^-17 |s_i|^2, in double precision; padded "impossible" spheres get -1e30 (never candidates).
__device__ __forceinline__ void build_sph_pairs_from_r2(const SceneView& sc, const float4* sph, float4* pairs)
{
    for (int p = threadIdx.x; 2 * p < sc.simdCount; p += blockDim.x)
    {
        const float4 a = sph[2 * p], b = sph[2 * p + 1];
        const double Sa = (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z, Sb = (double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z;
        const double m = 7.62939453125e-6;      // 2^-17
        // round the constant UP (towards the permissive side): the margin must not shrink by the rounding of -K
        const float na = __double2float_ru(Sa * m - (Sa - (double)a.w)), nbv = __double2float_ru(Sb * m - (Sb - (double)b.w));
        pairs[2 * p] = make_float4(a.x, b.x, a.y, b.y);
        pairs[2 * p + 1] = make_float4(a.z, b.z, 2 * p < sc.count ? na : -1.0e30f, 2 * p + 1 < sc.count ? nbv : -1.0e30f);
    }
}

// Pair array for FastHitterK2 from the {s, K} array build_sphK() left in place: pair p = spheres 2p, 2p+1 as
// {x0, x1, y0, y1} {z0, z1, -K0, -K1}.
__device__ __forceinline__ void build_sph_pairs(const SceneView& sc, const float4* sphK, float4* pairs)
{
    for (int p = threadIdx.x; 2 * p < sc.simdCount; p += blockDim.x)
    {
        const float4 a = sphK[2 * p], b = sphK[2 * p + 1];
        pairs[2 * p] = make_float4(a.x, b.x, a.y, b.y);
        pairs[2 * p + 1] = make_float4(a.z, b.z, -a.w, -b.w);
    }
}

// One iteration of the per-lane path state machine shared by the queue kernels: intersect the lane's current ray
// (path or shadow) against all spheres, then shade. Returns true when the lane's path has ended (st.col is final).
template <class Hitter>
__device__ __forceinline__ bool path_step(const SceneView& sc, QPath& st, unsigned& rc, const Hitter& hitter)
{
    // ---- intersect
    float t = TPT_MAX_T;
    int id = -1;
    if (st.active) { id = hitter.hit(sc, st.o, st.d, TPT_MIN_T, TPT_MAX_T, t); ++rc; }

    // ---- shade
    bool wantLight = false;
    int lightFrom = 0;
    bool finished = false;
    if (st.active)
    {
        if (st.kind == 0)
        {
            if (id < 0) { st.col = st.col + st.thr * sky(st.d, sc); finished = true; }
            else
            {
                Q4 s = ld_sph(sc, id);
                V3 pos = st.o + st.d * t;
                V3 normal = (pos - v3(s.x, s.y, s.z)) * sc.invRadius[id];
                const int mid = id < sc.count ? id : sc.count;
                Mat mat = load_mat(sc, mid);
                // the compiler turns a three-way test of mat.type into a jump table (LDC from the constant bank + BRX: two
                // long-latency steps in front of every shading branch); keeping the second test opaque keeps it two branches
                int mtype = mat.type;
                if (st.depth >= TPT_MAX_DEPTH) { st.col = st.col + st.thr * mat.emissive; finished = true; }
                else if (mtype == kLambert)
                {
                    if (st.doMaterialE) st.col = st.col + st.thr * mat.emissive;
                    V3 target = normal + RandomUnitVector<false>(st.rng);
                    st.nextDir = M<false>::normalize(target);
                    st.thrAlb = st.thr * mat.albedo;
                    st.albedo = mat.albedo;
                    st.nl = dot(normal, st.d) < 0.0f ? normal : neg(normal);
                    st.mid = mid;
                    st.o = pos;
                    wantLight = true; lightFrom = 0;
                }
                else if (opaque_int(mtype) == kMetal)
                {
                    // Test.cpp:137-150; with roughness == 0 the unit-sphere sample has zero weight, so the fast
                    // mode skips drawing it (the exact mode must draw it: it advances the shared RNG stream)
                    V3 refl = reflect(st.d, normal);
                    if (mat.roughness != 0.0f) refl = refl + mat.roughness * fast_in_unit_sphere(st.rng);
                    V3 outDir = M<false>::normalize(refl);
                    if (dot(outDir, normal) > 0.0f)
                    {
                        if (st.doMaterialE) st.col = st.col + st.thr * mat.emissive;
                        st.doMaterialE = true;
                        st.thr = st.thr * mat.albedo;
                        st.o = pos; st.d = outDir; ++st.depth;
                    }
                    else { st.col = st.col + st.thr * mat.emissive; finished = true; }
                }
                else
                {
                    V3 att, outDir;
                    bool ok = scatter_specular<false>(mat, st.d, pos, normal, st.rng, att, outDir);
                    if (!ok) { st.col = st.col + st.thr * mat.emissive; finished = true; }
                    else
                    {
                        if (st.doMaterialE) st.col = st.col + st.thr * mat.emissive;
                        st.doMaterialE = true;
                        st.thr = st.thr * att;
                        st.o = pos; st.d = outDir; ++st.depth;
                    }
                }
            }
        }
        else
        {
            const int j = st.kind - 1;
            if (id == sc.lights[j].id) st.col = st.col + st.pend;
            wantLight = true; lightFrom = j + 1;
        }
    }
    if (wantLight)
    {
        int j = lightFrom;
        while (j < sc.nLights && sc.lights[j].id == st.mid) ++j;
        if (j < sc.nLights)
        {
            const LightRec Lr = sc.lights[j];
            V3 scn = v3(Lr.cx, Lr.cy, Lr.cz);
            V3 pc = scn - st.o;
            float d2 = dot(pc, pc);
            float inv = __frsqrt_rn(d2);          // d2 is a squared distance between distinct spheres: never denormal
            V3 sw = pc * inv;
            // any orthonormal (su, sv) around sw gives the same cone-sample distribution (phi is uniform): use the
            // branch-free basis of Duff et al. 2017 instead of normalize(cross(up, sw)), cross(sw, su) (Test.cpp:108-109)
            const float sgn = copysignf(1.0f, sw.z);
            const float oa = __fdividef(-1.0f, sgn + sw.z);      // |sgn + sw.z| in [1, 2]: MUFU.RCP + FMUL instead of the IEEE sequence
            const float ob = sw.x * sw.y * oa;
            V3 su = v3(1.0f + sgn * sw.x * sw.x * oa, sgn * ob, -sgn * sw.x);
            V3 sv = v3(ob, sgn + sw.y * sw.y * oa, -sw.y);
            float cosAMax = M<false>::sqrt_(1.0f - Lr.radius * Lr.radius * inv * inv);
            float eps1 = RandomFloat01(st.rng), eps2 = RandomFloat01(st.rng);
            float cosA = 1.0f - eps1 + eps1 * cosAMax;
            float sinA = M<false>::sqrt_(1.0f - cosA * cosA);
            float phi = 2.0f * TPT_PI * eps2;
            float sp, cp;
            __sincosf(phi, &sp, &cp);
            V3 l = su * (cp * sinA) + sv * (sp * sinA) + sw * cosA;
            float omega = 2.0f * TPT_PI * (1.0f - cosAMax);
            float dl = dot(l, st.nl);
            float m = (0.0f < dl) ? dl : 0.0f;
            st.pend = st.thr * ((st.albedo * v3(Lr.ex, Lr.ey, Lr.ez)) * (m * omega * (1.0f / TPT_PI)));
            st.d = l;
            st.kind = 1 + j;
        }
        else
        {
            st.d = st.nextDir;
            st.thr = st.thrAlb;
            st.kind = 0;
            st.doMaterialE = false;
            ++st.depth;
        }
    }
    return finished;
}

// Diagnostic build (-DTPT_TRACE_WARPS=1, tools/warp_trace.py): every warp of k_fast_queue records global-timer stamps (entry,
// first slab, queue exhausted, exit), its trip counts and the lane-slots it used after the queue ran dry: the kernel's
// end effects (fill, drain, stragglers) measured instead of guessed. Not compiled into the shipped library.
#ifndef TPT_TRACE_WARPS
#define TPT_TRACE_WARPS 0
#endif
#if TPT_TRACE_WARPS
__device__ unsigned long long* g_warpTrace = nullptr;
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

// KFORM 0: reference-form sweep, 1: expanded form, 2: expanded form on packed pairs (FFMA2), 3: packed-pair pass 1 made
// conservative + reference-form pass 2 (FastHitterK2C: any scene). THREADS: 128 (6 CTAs/SM) for scenes whose staged
// geometry is small; one big CTA per SM when the sphere arrays fill most of an SM's shared memory.
// Dynamic shared memory: [staged blob][pair array (KFORM 2/3)][per-warp camera-ray buffers, 2 KB per warp at raysOffset].
template <int THREADS, int MINB, int KFORM>
__global__ void __launch_bounds__(THREADS, MINB)
k_fast_queue(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights,
             uint32_t stagedBytes, uint32_t numSlabs, uint32_t S, unsigned int* __restrict__ bandDone, uint32_t mtilesPerBand,
             uint32_t raysOffset)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ float sW[kMaxFramesPerDraw];
    // camera rays of the warp's current slab, generated 4 per lane in one convergent burst when the slab is fetched:
    // {origin.xyz, rng state} {direction.xyz, pixel offset}; regeneration then only pops an entry.
    float4 (*sRays)[kSlabPix][2] = reinterpret_cast<float4 (*)[kSlabPix][2]>(smem + raysOffset);
#if TPT_TRACE_WARPS
    const unsigned long long trEntry = gtimer();
    unsigned long long trReady = 0, trDry = 0;
    unsigned trTrips = 0, trDryTrips = 0, trDryLanes = 0;
#endif
    stage_blob(smem, blob, stagedBytes, &bar);
    if (threadIdx.x == 0) { float wp; blend_weights(p, sW, wp); }
    // the 128-thread instances are only launched for scenes whose whole blob is staged (launch_fast): their view points into
    // shared memory for every section, so materials, lights and 1/r load with LDS instead of generic LD
    SceneView sc = make_view<THREADS == 128>(smem, blob, L, stagedBytes, count, nLights);
    float4* sphK = reinterpret_cast<float4*>(smem + L.offSph);
    if (KFORM == 1 || KFORM == 2) build_sphK(sc, sphK);
    __syncthreads();
    float4* pairs = reinterpret_cast<float4*>(smem + ((stagedBytes + 127u) & ~127u));     // KFORM 2/3: behind the staged blob
    if (KFORM == 2) { build_sph_pairs(sc, sphK, pairs); __syncthreads(); }
    if (KFORM == 3) { build_sph_pairs_from_r2(sc, sphK, pairs); __syncthreads(); }
    FastHitterK hitK; hitK.sphK = sc.sphShared; hitK.simdCount = sc.simdCount;
    FastHitterK2 hitK2; hitK2.sphK = sc.sphShared; hitK2.sphP = smem_u32(pairs); hitK2.simdCount = sc.simdCount;
    FastHitterK2C hitK2C; hitK2C.sph = sc.sphShared; hitK2C.sphP = smem_u32(pairs); hitK2C.simdCount = sc.simdCount;
    // the sweep's loads are plain (schedulable) asm: make their address opaque AFTER the barrier so that none of them can
    // be hoisted above the in-place {s, r^2} -> {s, K} rewrite / the construction of the pair array
    asm volatile("" : "+r"(hitK.sphK), "+r"(hitK2.sphK), "+r"(hitK2.sphP), "+r"(hitK2C.sph), "+r"(hitK2C.sphP));
    SerialHitter<false> hitS;
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const uint32_t regionPix = (uint32_t)((long long)p.numRows * p.width);
    const float invSpp = 1.0f / (float)p.spp;
    
    unsigned rc = 0;

    // warp-uniform slab cursor
    uint32_t slabCur = 0, slabEnd = 0;
    int slabX0 = 0, slabRi0 = 0;
    uint32_t slabSample = 0, slabFrame = 0;
    float slabW = 0.0f;
    bool exhausted = false;
    // progress reporting for host-buffer draws (bandDone != nullptr): slabs are dealt in pixel order, so the image
    // completes band by band; every lane counts the paths it finished per band and publishes the count (after a
    // __threadfence so its reductions are visible first) when it moves on to the next band. The host's copy stream
    // waits on these counters (cuStreamWaitValue32) and starts the D2H of a band while later bands are still traced.
    // (Measured: whichever band is dealt first completes at ~50 % of the kernel, not at 1/bands — warps progress
    // unevenly and a few keep their first slab for a long time — so about half of the copy overlaps; dealing slabs from
    // both ends of the image by hardware warp slot, and cheap-rows-first ordering, did not change that.)
    uint32_t slabBand = 0;
    int curBand = -1;
    uint32_t doneCnt = 0, myBand = 0;

    QPath st;
    st.active = false;
    for (;;)
    {
        // ---- regeneration
        unsigned need = __ballot_sync(0xffffffffu, !st.active);
        while (need && !exhausted)
        {
            if (slabCur >= slabEnd)
            {
                uint32_t slab = 0;
                if (lane == 0) slab = atomicAdd(p.workCounter, 1u);
                slab = __shfl_sync(0xffffffffu, slab, 0);
                if (slab >= numSlabs) { exhausted = true; break; }
                uint32_t mtile = slab / S;
                const uint32_t s = slab - mtile * S;
                // progress mode: walk the image from its LAST macro-tile to its first. Row 0 is the bottom of the image
                // (ground, light sampling: the expensive pixels); starting with the cheap upper rows makes pixels
                // complete early at a high rate, so most of the D2H runs under the expensive rows' tracing.
                if (bandDone) mtile = (numSlabs / S) - 1u - mtile;
                const uint32_t pix0 = mtile * kSlabPix;
                slabEnd = regionPix - pix0 < (uint32_t)kSlabPix ? regionPix - pix0 : (uint32_t)kSlabPix;
                slabCur = 0;
                slabRi0 = (int)(pix0 / (uint32_t)p.width);
                slabX0 = (int)(pix0 - (uint32_t)slabRi0 * (uint32_t)p.width);
                const uint32_t fi = s / (uint32_t)p.spp;
                slabSample = s - fi * (uint32_t)p.spp;
                slabFrame = (uint32_t)p.frame0 + fi;
                slabW = invSpp * sW[fi];
                slabBand = bandDone ? mtile / mtilesPerBand : 0u;
                __syncwarp();       // every lane has popped what it needed from the previous slab
                generate_slab_rays(p, sRays[threadIdx.x >> 5], lane, slabEnd, slabX0, slabRi0, slabSample, slabFrame, -1);
                __syncwarp();
            }
            const uint32_t avail = slabEnd - slabCur;
            const uint32_t rank = (uint32_t)__popc(need & ltMask);
            if (!st.active && rank < avail)
            {
                const float4 e0 = sRays[threadIdx.x >> 5][slabCur + rank][0];
                const float4 e1 = sRays[threadIdx.x >> 5][slabCur + rank][1];
                st.o = v3(e0.x, e0.y, e0.z); st.d = v3(e1.x, e1.y, e1.z);
                st.rng = __float_as_uint(e0.w);
                st.pixOff = __float_as_uint(e1.w);
                st.thr = v3(1, 1, 1); st.col = v3(0, 0, 0);
                st.weight = slabW;
                myBand = slabBand;
                st.kind = 0; st.depth = 0; st.doMaterialE = true; st.active = true;
            }
            const uint32_t n = (uint32_t)__popc(need);
            slabCur += n < avail ? n : avail;
            need = __ballot_sync(0xffffffffu, !st.active);
        }
        if (!__any_sync(0xffffffffu, st.active)) break;
#if TPT_TRACE_WARPS
        if (!trReady) trReady = gtimer();
        ++trTrips;
        if (exhausted) { if (!trDry) trDry = gtimer(); ++trDryTrips; trDryLanes += (unsigned)__popc(__ballot_sync(0xffffffffu, st.active)); }
#endif

        const bool finished = KFORM == 3 ? path_step(sc, st, rc, hitK2C) : KFORM == 2 ? path_step(sc, st, rc, hitK2) : (KFORM == 1 ? path_step(sc, st, rc, hitK) : path_step(sc, st, rc, hitS));
        if (finished)
        {
            red_add_f4(p.image + (size_t)st.pixOff * 4, st.col.x * st.weight, st.col.y * st.weight, st.col.z * st.weight);
            st.active = false;
            if (bandDone)
            {
                if ((int)myBand != curBand)
                {
                    if (doneCnt) { __threadfence(); atomicAdd(bandDone + curBand, doneCnt); }
                    curBand = (int)myBand; doneCnt = 0;
                }
                ++doneCnt;
            }
        }
    }
    if (bandDone && doneCnt) { __threadfence(); atomicAdd(bandDone + curBand, doneCnt); }
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(p.rayCounter, (unsigned long long)rc);
#if TPT_TRACE_WARPS
    if (lane == 0 && g_warpTrace)
    {
        unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        unsigned long long* t = g_warpTrace + ((size_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * 8;
        t[0] = trEntry; t[1] = trReady; t[2] = trDry; t[3] = gtimer(); t[4] = trTrips; t[5] = trDryTrips; t[6] = trDryLanes; t[7] = smid;
    }
#endif
}

// ---- variant 8 ------------------------------------------------------------------------------------------------
// "warp-owned pixel groups": the warp-level dealing and path state machine of variant 3, but the unit a warp pulls from
// the global counter is a GROUP of kGroupPix consecutive pixels with ALL their samples of this draw. The group's radiance
// is accumulated in the warp's own shared-memory slot; when its last path ends, 16 lanes write the 16 finished float4
// pixels with ONE coalesced 128-bit store each (256 contiguous bytes) — straight into the caller's buffer: local HBM,
// a peer GPU's HBM over NVLink (multi-GPU write-out), or page-locked HOST memory over PCIe (host-buffer draws finish
// with the kernel: no staging image, no device-to-host copy). No prepare kernel, no L2 reductions, no block barriers
// after the prologue. Because of ray regeneration a warp has paths of several groups in flight: kGroupOpen slots per
// warp; a new group is only opened in a slot whose previous group has been written out.
constexpr int kGroupPix = 16;
constexpr int kGroupOpen = 8;

// Camera rays of one chunk of a group: `ns` sample indices [s0, s0+ns) x npix pixels, entry q = sl*npix + px, into the warp's
// shared buffer: {origin.xyz, rng} {direction.xyz, px | slot << 4 | frame-in-draw << 8}.
__device__ __forceinline__ void generate_group_rays(const DrawParams& p, float4 (*rays)[2], int lane, uint32_t npix, uint32_t ns,
                                                    uint32_t pix0, uint32_t s0, uint32_t slot)
{
    for (uint32_t q = (uint32_t)lane; q < npix * ns; q += 32)
    {
        const uint32_t sl = npix == (uint32_t)kGroupPix ? q >> 4 : q / npix;
        const uint32_t px = q - sl * npix;
        const uint32_t gp = pix0 + px;
        const int ri = (int)(gp / (uint32_t)p.width), x = (int)(gp - (uint32_t)ri * (uint32_t)p.width);
        const int y = p.row0 + ri * p.rowStep;
        const uint32_t s = s0 + sl, fi = s / (uint32_t)p.spp, ss = s - fi * (uint32_t)p.spp;
        uint32_t rng = pixel_seed((uint32_t)(y * p.width + x) * (uint32_t)p.spp + ss, (uint32_t)p.frame0 + fi);
        float u = ((float)x + RandomFloat01(rng)) * p.invWidth;
        float v = ((float)y + RandomFloat01(rng)) * p.invHeight;
        Ray r = fast_get_ray(p.cam, u, v, rng);
        rays[q][0] = make_float4(r.orig.x, r.orig.y, r.orig.z, __uint_as_float(rng));
        rays[q][1] = make_float4(r.dir.x, r.dir.y, r.dir.z, __uint_as_float(px | (slot << 4) | (fi << 8)));
    }
}

template <int MINB, int KFORM, bool ALLS>
__global__ void __launch_bounds__(kQueueThreads, MINB)
k_fast_group(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights,
             uint32_t stagedBytes, uint32_t numGroups, uint32_t S, float wPrev)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ float sW[kMaxFramesPerDraw];
    __shared__ float4 sRays[kQueueThreads / 32][kSlabPix][2];
    __shared__ float4 sAcc[kQueueThreads / 32][kGroupOpen][kGroupPix];
    __shared__ int sRemain[kQueueThreads / 32][kGroupOpen];
    __shared__ uint32_t sGrpPix0[kQueueThreads / 32][kGroupOpen];
    stage_blob(smem, blob, stagedBytes, &bar);
    if (threadIdx.x == 0) { float wp; blend_weights(p, sW, wp); }
    for (int i = threadIdx.x; i < (kQueueThreads / 32) * kGroupOpen * kGroupPix; i += kQueueThreads) (&sAcc[0][0][0])[i] = make_float4(0, 0, 0, 0);
    if (threadIdx.x < (kQueueThreads / 32) * kGroupOpen) (&sRemain[0][0])[threadIdx.x] = 0;
    SceneView sc = make_view<ALLS>(smem, blob, L, stagedBytes, count, nLights);     // ALLS: whole blob staged -> LDS for every section
    float4* sphK = reinterpret_cast<float4*>(smem + L.offSph);
    if (KFORM) build_sphK(sc, sphK);
    __syncthreads();
    float4* pairs = reinterpret_cast<float4*>(smem + ((stagedBytes + 127u) & ~127u));
    if (KFORM == 2) { build_sph_pairs(sc, sphK, pairs); __syncthreads(); }
    FastHitterK hitK; hitK.sphK = sc.sphShared; hitK.simdCount = sc.simdCount;
    FastHitterK2 hitK2; hitK2.sphK = sc.sphShared; hitK2.sphP = smem_u32(pairs); hitK2.simdCount = sc.simdCount;
    asm volatile("" : "+r"(hitK.sphK), "+r"(hitK2.sphK), "+r"(hitK2.sphP));     // see k_fast_queue
    SerialHitter<false> hitS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    const uint32_t regionPix = (uint32_t)((long long)p.numRows * p.width);
    const float invSpp = 1.0f / (float)p.spp;
    unsigned rc = 0;

    // warp-uniform cursors: current chunk of rays in sRays, current group
    uint32_t chunkCur = 0, chunkEnd = 0;
    uint32_t sNext = S;                     // next sample index of the current group to generate (== S: group exhausted)
    uint32_t grpPix0 = 0, grpNpix = 0, curSlot = 0, groupsOpened = 0;
    bool exhausted = false;

    QPath st;
    st.active = false;
    for (;;)
    {
        // ---- regeneration
        unsigned need = __ballot_sync(0xffffffffu, !st.active);
        while (need && !exhausted)
        {
            if (chunkCur >= chunkEnd)
            {
                if (sNext >= S)
                {
                    const uint32_t slot = groupsOpened % (uint32_t)kGroupOpen;
                    if (sRemain[warp][slot] != 0) break;        // that slot's group still has paths in flight: trace them first
                    uint32_t g = 0;
                    if (lane == 0) g = atomicAdd(p.workCounter, 1u);
                    g = __shfl_sync(0xffffffffu, g, 0);
                    if (g >= numGroups) { exhausted = true; break; }
                    grpPix0 = g * (uint32_t)kGroupPix;
                    grpNpix = regionPix - grpPix0 < (uint32_t)kGroupPix ? regionPix - grpPix0 : (uint32_t)kGroupPix;
                    curSlot = slot; ++groupsOpened; sNext = 0;
                    if (lane == 0) { sRemain[warp][slot] = (int)(grpNpix * S); sGrpPix0[warp][slot] = grpPix0; }
                }
                const uint32_t ns = S - sNext < 4u ? S - sNext : 4u;
                __syncwarp();       // every lane has popped what it needed from the previous chunk
                generate_group_rays(p, sRays[warp], lane, grpNpix, ns, grpPix0, sNext, curSlot);
                __syncwarp();
                chunkCur = 0; chunkEnd = grpNpix * ns; sNext += ns;
            }
            const uint32_t avail = chunkEnd - chunkCur;
            const uint32_t rank = (uint32_t)__popc(need & ltMask);
            if (!st.active && rank < avail)
            {
                const float4 e0 = sRays[warp][chunkCur + rank][0];
                const float4 e1 = sRays[warp][chunkCur + rank][1];
                st.o = v3(e0.x, e0.y, e0.z); st.d = v3(e1.x, e1.y, e1.z);
                st.rng = __float_as_uint(e0.w);
                const uint32_t tag = __float_as_uint(e1.w);
                st.pixOff = tag & 0xffu;                         // px | slot << 4
                st.weight = invSpp * sW[tag >> 8];
                st.thr = v3(1, 1, 1); st.col = v3(0, 0, 0);
                st.kind = 0; st.depth = 0; st.doMaterialE = true; st.active = true;
            }
            const uint32_t n = (uint32_t)__popc(need);
            chunkCur += n < avail ? n : avail;
            need = __ballot_sync(0xffffffffu, !st.active);
        }
        if (!__any_sync(0xffffffffu, st.active)) break;

        const bool finished = KFORM == 2 ? path_step(sc, st, rc, hitK2) : (KFORM == 1 ? path_step(sc, st, rc, hitK) : path_step(sc, st, rc, hitS));
        bool lastOfGroup = false;
        if (finished)
        {
            float* a = reinterpret_cast<float*>(&sAcc[warp][st.pixOff >> 4][st.pixOff & 15u]);
            atomicAdd(a + 0, st.col.x * st.weight);
            atomicAdd(a + 1, st.col.y * st.weight);
            atomicAdd(a + 2, st.col.z * st.weight);
            st.active = false;
            lastOfGroup = atomicSub(&sRemain[warp][st.pixOff >> 4], 1) == 1;
        }
        // ---- a group completed: 16 lanes write its 16 pixels, one coalesced 128-bit store each
        unsigned trig = __ballot_sync(0xffffffffu, lastOfGroup);
        while (trig)
        {
            const int src = __ffs(trig) - 1;
            trig &= trig - 1;
            const uint32_t slot = __shfl_sync(0xffffffffu, st.pixOff >> 4, src);
            const uint32_t pix0 = sGrpPix0[warp][slot];
            const uint32_t npix = regionPix - pix0 < (uint32_t)kGroupPix ? regionPix - pix0 : (uint32_t)kGroupPix;
            if ((uint32_t)lane < npix)
            {
                const uint32_t gp = pix0 + (uint32_t)lane;
                const int ri = (int)(gp / (uint32_t)p.width), x = (int)(gp - (uint32_t)ri * (uint32_t)p.width);
                const int y = p.row0 + ri * p.rowStep;
                float* px = p.image + ((size_t)(p.packed ? ri : y) * p.width + x) * 4;
                const float4 a = sAcc[warp][slot][lane];
                st_stream_f4(px, blend_prev(p, px, wPrev, a.x, a.y, a.z));
                sAcc[warp][slot][lane] = make_float4(0, 0, 0, 0);
            }
            __syncwarp();
        }
    }
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(p.rayCounter, (unsigned long long)rc);
}

// ---- variant 5 ------------------------------------------------------------------------------------------------
// "tile queue": the warp-level slab dealing and path state machine of variant 3, but a CTA owns a tile of kTileQPix
// pixels for ALL its samples: radiance is accumulated in shared memory and the finished tile is written ONCE with
// coalesced 128-bit stores (st.global.L1::no_allocate.v4, + 128-bit loads of `prev` when it has weight). The store
// target may be another GPU's memory (CUDA IPC mapping): this is the fused render + gather of the multi-GPU path —
// pixels leave over NVLink tile by tile while the other tiles are still being traced. One block barrier per tile; with
// S = spp x frames >= 16 the per-tile tail is < 1 %, at S = 4 variant 3 is the better choice.
constexpr int kTileQPix = 1024;

template <int MINB, int KFORM>
__global__ void __launch_bounds__(kQueueThreads, MINB)
k_fast_tileq(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights,
             uint32_t stagedBytes, uint32_t numTiles, uint32_t S, float wPrev, uint32_t tileQPix)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ float sW[kMaxFramesPerDraw];
    __shared__ float sAcc[kTileQPix * 3];
    __shared__ float4 sRays[kQueueThreads / 32][kSlabPix][2];   // see k_fast_queue
    __shared__ uint32_t sTile, sSlab;
    stage_blob(smem, blob, stagedBytes, &bar);
    if (threadIdx.x == 0) { float wp; blend_weights(p, sW, wp); }
    for (int i = threadIdx.x; i < kTileQPix * 3; i += kQueueThreads) sAcc[i] = 0.0f;
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);
    float4* sphK = reinterpret_cast<float4*>(smem + L.offSph);
    if (KFORM) build_sphK(sc, sphK);
    __syncthreads();
    float4* pairs = reinterpret_cast<float4*>(smem + ((stagedBytes + 127u) & ~127u));
    if (KFORM == 2) { build_sph_pairs(sc, sphK, pairs); __syncthreads(); }
    FastHitterK hitK; hitK.sphK = sc.sphShared; hitK.simdCount = sc.simdCount;
    FastHitterK2 hitK2; hitK2.sphK = sc.sphShared; hitK2.sphP = smem_u32(pairs); hitK2.simdCount = sc.simdCount;
    asm volatile("" : "+r"(hitK.sphK), "+r"(hitK2.sphK), "+r"(hitK2.sphP));     // see k_fast_queue
    SerialHitter<false> hitS;
    const int lane = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const uint32_t regionPix = (uint32_t)((long long)p.numRows * p.width);
    const float invSpp = 1.0f / (float)p.spp;
    unsigned rc = 0;

    for (;;)
    {
        __syncthreads();
        if (threadIdx.x == 0) { sTile = atomicAdd(p.workCounter, 1u); sSlab = 0; }
        __syncthreads();
        const uint32_t tile = sTile;
        if (tile >= numTiles) break;
        const uint32_t tilePix0 = tile * tileQPix;
        const uint32_t tilePix = regionPix - tilePix0 < tileQPix ? regionPix - tilePix0 : tileQPix;
        const uint32_t slabsPerSample = (tilePix + kSlabPix - 1) / kSlabPix;
        const uint32_t tileSlabs = slabsPerSample * S;

        uint32_t slabCur = 0, slabEnd = 0, slabQ0 = 0;
        int slabX0 = 0, slabRi0 = 0;
        uint32_t slabSample = 0, slabFrame = 0;
        float slabW = 0.0f;
        bool exhausted = false;
        QPath st;
        st.active = false;
        for (;;)
        {
            unsigned need = __ballot_sync(0xffffffffu, !st.active);
            while (need && !exhausted)
            {
                if (slabCur >= slabEnd)
                {
                    uint32_t slab = 0;
                    if (lane == 0) slab = atomicAdd(&sSlab, 1u);
                    slab = __shfl_sync(0xffffffffu, slab, 0);
                    if (slab >= tileSlabs) { exhausted = true; break; }
                    const uint32_t s = slab / slabsPerSample, g = slab - s * slabsPerSample;   // sample-major inside the tile
                    slabQ0 = g * kSlabPix;
                    slabEnd = tilePix - slabQ0 < (uint32_t)kSlabPix ? tilePix - slabQ0 : (uint32_t)kSlabPix;
                    slabCur = 0;
                    const uint32_t pix0 = tilePix0 + slabQ0;
                    slabRi0 = (int)(pix0 / (uint32_t)p.width);
                    slabX0 = (int)(pix0 - (uint32_t)slabRi0 * (uint32_t)p.width);
                    const uint32_t fi = s / (uint32_t)p.spp;
                    slabSample = s - fi * (uint32_t)p.spp;
                    slabFrame = (uint32_t)p.frame0 + fi;
                    slabW = invSpp * sW[fi];
                    __syncwarp();
                    generate_slab_rays(p, sRays[threadIdx.x >> 5], lane, slabEnd, slabX0, slabRi0, slabSample, slabFrame, (int)slabQ0);
                    __syncwarp();
                }
                const uint32_t avail = slabEnd - slabCur;
                const uint32_t rank = (uint32_t)__popc(need & ltMask);
                if (!st.active && rank < avail)
                {
                    const float4 e0 = sRays[threadIdx.x >> 5][slabCur + rank][0];
                    const float4 e1 = sRays[threadIdx.x >> 5][slabCur + rank][1];
                    st.o = v3(e0.x, e0.y, e0.z); st.d = v3(e1.x, e1.y, e1.z);
                    st.rng = __float_as_uint(e0.w);
                    st.pixOff = __float_as_uint(e1.w);      // pixel index inside the tile
                    st.thr = v3(1, 1, 1); st.col = v3(0, 0, 0);
                    st.weight = slabW;
                    st.kind = 0; st.depth = 0; st.doMaterialE = true; st.active = true;
                }
                const uint32_t n = (uint32_t)__popc(need);
                slabCur += n < avail ? n : avail;
                need = __ballot_sync(0xffffffffu, !st.active);
            }
            if (!__any_sync(0xffffffffu, st.active)) break;
            if (KFORM == 2 ? path_step(sc, st, rc, hitK2) : (KFORM == 1 ? path_step(sc, st, rc, hitK) : path_step(sc, st, rc, hitS)))
            {
                atomicAdd(&sAcc[st.pixOff * 3 + 0], st.col.x * st.weight);
                atomicAdd(&sAcc[st.pixOff * 3 + 1], st.col.y * st.weight);
                atomicAdd(&sAcc[st.pixOff * 3 + 2], st.col.z * st.weight);
                st.active = false;
            }
        }

        // ---- tile finished: coalesced 128-bit write-out (possibly into a peer GPU's memory)
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < tilePix; i += kQueueThreads)
        {
            const uint32_t gp = tilePix0 + i;
            const int ri = (int)(gp / (uint32_t)p.width), x = (int)(gp - (uint32_t)ri * (uint32_t)p.width);
            const int y = p.row0 + ri * p.rowStep;
            float* px = p.image + ((size_t)(p.packed ? ri : y) * p.width + x) * 4;
            st_stream_f4(px, blend_prev(p, px, wPrev, sAcc[i * 3 + 0], sAcc[i * 3 + 1], sAcc[i * 3 + 2]));
            sAcc[i * 3 + 0] = 0.0f; sAcc[i * 3 + 1] = 0.0f; sAcc[i * 3 + 2] = 0.0f;
        }
    }
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(p.rayCounter, (unsigned long long)rc);
}

// ---- variant 6 ------------------------------------------------------------------------------------------------
// "block wavefront": the persistent design of the north star taken literally. A CTA keeps kWaveSlots path slots in
// SHARED memory (11 float4 per slot) and runs every iteration in three phases separated by block barriers:
//   sweep     thread i regenerates slot i if it is free (next path of the CTA's chunk) and intersects slot i's
//             ray (path or shadow) against all spheres -> every lane has a ray, no divergence outside pass 2;
//   sort      the slot is classified {miss, lambert, metal, dielectric, shadow-return, terminal} and appended to
//             that type's index list with warp __ballot_sync + prefix popcount (one shared atomic per warp per type);
//   scatter   the lists are cut into 32-entry chunks and dealt to the warps: every warp shades entries of ONE
//             material type (Scatter(), Test.cpp:83-193, without divergence); Lambert vertices and returning
//             shadow rays append themselves to the light list, processed the same way in a third phase
//             (explicit light sampling, Test.cpp:96-133).
// Radiance leaves through the same 128-bit vector reductions as variant 3.
constexpr int kWaveThreads = 256;
constexpr int kWaveSlots = 256;        // one slot per thread
constexpr int kWaveFields = 11;
constexpr int kWaveChunk = 4096;       // paths taken from the global counter at a time
enum { WF_O = 0, WF_D, WF_THR, WF_COL, WF_NEXT, WF_THRALB, WF_NL, WF_ALB, WF_PEND, WF_POS, WF_NRM };
enum { WT_MISS = 0, WT_LAMBERT, WT_METAL, WT_DIEL, WT_SHADOW, WT_TERMINAL, WT_COUNT };
constexpr int kKindFree = 255;

struct WaveArgs
{
    uint32_t totalPaths, S;
    FastDiv divS, divW, divSpp;
};

template <int MINB>
__global__ void __launch_bounds__(kWaveThreads, MINB)
k_fast_wave(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights,
            uint32_t stagedBytes, uint32_t poolOffset, WaveArgs wa)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ float sW[kMaxFramesPerDraw];
    __shared__ uint16_t sList[WT_COUNT][kWaveSlots];
    __shared__ uint16_t sLight[kWaveSlots];
    __shared__ int sCount[WT_COUNT];
    __shared__ int sLightCount;
    __shared__ uint32_t sNextPath, sEndPath;
    __shared__ int sExhausted, sLive;
    stage_blob(smem, blob, stagedBytes, &bar);
    float4* pool = reinterpret_cast<float4*>(smem + poolOffset);
#define WFLD(f, slot) pool[(f) * kWaveSlots + (slot)]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    if (tid == 0)
    {
        float wp; blend_weights(p, sW, wp);
        sNextPath = 0; sEndPath = 0; sExhausted = 0; sLive = 0; sLightCount = 0;
        for (int t = 0; t < WT_COUNT; ++t) sCount[t] = 0;
    }
    WFLD(WF_D, tid) = make_float4(0, 0, 0, __int_as_float(kKindFree));
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);
    const float invSpp = 1.0f / (float)p.spp;
    SerialHitter<false> hitter;
    unsigned rc = 0;
    __syncthreads();

    for (;;)
    {
        // ---- list reset + chunk refill (thread 0 only, between the previous iteration's last barrier and this one)
        if (tid == 0)
        {
            sLightCount = 0;
            for (int t = 0; t < WT_COUNT; ++t) sCount[t] = 0;
            if (!sExhausted && sNextPath >= sEndPath)
            {
                const uint32_t base = atomicAdd(p.workCounter, (unsigned)kWaveChunk);
                if (base >= wa.totalPaths) sExhausted = 1;
                else { sNextPath = base; sEndPath = base + kWaveChunk < wa.totalPaths ? base + kWaveChunk : wa.totalPaths; }
            }
        }
        __syncthreads();

        // ---- phase 1: regenerate own slot if free, then sweep
        float4 fd = WFLD(WF_D, tid);
        int misc = __float_as_int(fd.w);
        int kind = misc & 0xff;
        {
            const unsigned need = __ballot_sync(0xffffffffu, kind == kKindFree);
            if (need)
            {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&sNextPath, (uint32_t)__popc(need));
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t idx = base + (uint32_t)__popc(need & ltMask);
                if (kind == kKindFree && idx < sEndPath)
                {
                    const uint32_t slab = idx >> 7, q = idx & 127u;
                    const uint32_t mtile = fdiv(slab, wa.divS), s = slab - mtile * wa.S;
                    uint32_t pix = mtile * 128u + q;
                    const uint32_t regionPix = (uint32_t)p.numRows * (uint32_t)p.width;
                    if (pix < regionPix)
                    {
                        const uint32_t ri = fdiv(pix, wa.divW), x = pix - ri * (uint32_t)p.width;
                        const uint32_t fi = fdiv(s, wa.divSpp), ss = s - fi * (uint32_t)p.spp;
                        const int y = p.row0 + (int)ri * p.rowStep;
                        uint32_t rng = pixel_seed((uint32_t)(y * p.width + (int)x) * (uint32_t)p.spp + ss, (uint32_t)p.frame0 + fi);
                        float u = ((float)x + RandomFloat01(rng)) * p.invWidth;
                        float v = ((float)y + RandomFloat01(rng)) * p.invHeight;
                        Ray r = fast_get_ray(p.cam, u, v, rng);
                        kind = 0; misc = 0 | (0 << 8) | (1 << 16);
                        fd = make_float4(r.dir.x, r.dir.y, r.dir.z, __int_as_float(misc));
                        WFLD(WF_O, tid) = make_float4(r.orig.x, r.orig.y, r.orig.z, 0.0f);
                        WFLD(WF_D, tid) = fd;
                        WFLD(WF_THR, tid) = make_float4(1, 1, 1, __uint_as_float(rng));
                        WFLD(WF_COL, tid) = make_float4(0, 0, 0, __uint_as_float((uint32_t)((p.packed ? (int)ri : y) * p.width + (int)x)));
                        WFLD(WF_NEXT, tid) = make_float4(0, 0, 0, invSpp * sW[fi]);
                    }
                }
            }
        }
        int type = -1;
        if (kind != kKindFree)
        {
            const float4 fo = WFLD(WF_O, tid);
            const V3 o = v3(fo.x, fo.y, fo.z), d = v3(fd.x, fd.y, fd.z);
            float t;
            const int id = hitter.hit(sc, o, d, TPT_MIN_T, TPT_MAX_T, t);
            ++rc;
            if (kind != 0) { type = WT_SHADOW; WFLD(WF_POS, tid).w = __int_as_float(id); }
            else if (id < 0) type = WT_MISS;
            else
            {
                const Q4 s = ld_sph(sc, id);
                const V3 pos = o + d * t;
                const V3 normal = (pos - v3(s.x, s.y, s.z)) * sc.invRadius[id];
                const int mid = id < sc.count ? id : sc.count;
                const int depth = (misc >> 8) & 0xff;
                const int mtype = f_as_i(sc.matA[mid].w);
                type = depth >= TPT_MAX_DEPTH ? WT_TERMINAL : (mtype == kLambert ? WT_LAMBERT : (mtype == kMetal ? WT_METAL : WT_DIEL));
                WFLD(WF_POS, tid) = make_float4(pos.x, pos.y, pos.z, __int_as_float(mid));
                WFLD(WF_NRM, tid) = make_float4(normal.x, normal.y, normal.z, 0.0f);
            }
        }
        // ---- phase 2: sort slots by type (ballot + prefix)
#pragma unroll
        for (int T = 0; T < WT_COUNT; ++T)
        {
            const unsigned m = __ballot_sync(0xffffffffu, type == T);
            if (m)
            {
                int base = 0;
                if (lane == 0) base = atomicAdd(&sCount[T], __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (type == T) sList[T][base + __popc(m & ltMask)] = (uint16_t)tid;
            }
        }
        __syncthreads();
        const int live = sCount[0] + sCount[1] + sCount[2] + sCount[3] + sCount[4] + sCount[5];
        if (live == 0 && sExhausted) break;

        // ---- phase 3: Scatter() per material type, 32-entry chunks dealt to the warps
        {
            int chunk = warp;               // chunk ids over the concatenation of the type lists
            int T = 0, first = 0;           // `first` = chunk id of type T's first chunk
            for (;;)
            {
                while (T < WT_COUNT && chunk >= first + ((sCount[T] + 31) >> 5)) { first += (sCount[T] + 31) >> 5; ++T; }
                if (T >= WT_COUNT) break;
                const int k = ((chunk - first) << 5) + lane;
                const bool on = k < sCount[T];
                const int slot = on ? sList[T][k] : 0;
                bool wantLight = false, finished = false;
                float4 fcol, fthr;
                if (on)
                {
                    fcol = WFLD(WF_COL, slot);
                    fthr = WFLD(WF_THR, slot);
                    float4 fdd = WFLD(WF_D, slot);
                    int m2 = __float_as_int(fdd.w);
                    V3 col = v3(fcol.x, fcol.y, fcol.z), thr = v3(fthr.x, fthr.y, fthr.z);
                    uint32_t rng = __float_as_uint(fthr.w);
                    const V3 d = v3(fdd.x, fdd.y, fdd.z);
                    const bool doMatE = (m2 >> 16) & 1;
                    if (T == WT_MISS) { col = col + thr * sky(d, sc); finished = true; }
                    else if (T == WT_SHADOW)
                    {
                        const int j = (m2 & 0xff) - 1;
                        const int hid = __float_as_int(WFLD(WF_POS, slot).w);
                        if (hid == sc.lights[j].id) { const float4 pe = WFLD(WF_PEND, slot); col = col + v3(pe.x, pe.y, pe.z); }
                        wantLight = true;   // next light index = j + 1, kept in kind
                    }
                    else
                    {
                        const float4 fp = WFLD(WF_POS, slot), fn = WFLD(WF_NRM, slot);
                        const V3 pos = v3(fp.x, fp.y, fp.z), normal = v3(fn.x, fn.y, fn.z);
                        const int mid = __float_as_int(fp.w);
                        const Mat mat = load_mat(sc, mid);
                        if (T == WT_TERMINAL) { col = col + thr * mat.emissive; finished = true; }
                        else if (T == WT_LAMBERT)
                        {
                            if (doMatE) col = col + thr * mat.emissive;
                            const V3 target = normal + RandomUnitVector<false>(rng);
                            const V3 nextDir = M<false>::normalize(target);
                            const V3 nl = dot(normal, d) < 0.0f ? normal : neg(normal);
                            const V3 thrAlb = thr * mat.albedo;
                            WFLD(WF_NEXT, slot) = make_float4(nextDir.x, nextDir.y, nextDir.z, WFLD(WF_NEXT, slot).w);
                            WFLD(WF_THRALB, slot) = make_float4(thrAlb.x, thrAlb.y, thrAlb.z, __int_as_float(mid));
                            WFLD(WF_NL, slot) = make_float4(nl.x, nl.y, nl.z, 0.0f);
                            WFLD(WF_ALB, slot) = make_float4(mat.albedo.x, mat.albedo.y, mat.albedo.z, 0.0f);
                            WFLD(WF_O, slot) = make_float4(pos.x, pos.y, pos.z, 0.0f);
                            m2 = (m2 & ~0xff) | 0;      // kind := "before light 0" (light phase advances it)
                            wantLight = true;
                        }
                        else
                        {
                            V3 att, outDir;
                            bool ok;
                            if (T == WT_METAL)
                            {
                                V3 refl = reflect(d, normal);
                                if (mat.roughness != 0.0f) refl = refl + mat.roughness * fast_in_unit_sphere(rng);
                                outDir = M<false>::normalize(refl);
                                att = mat.albedo;
                                ok = dot(outDir, normal) > 0.0f;
                            }
                            else ok = scatter_specular<false>(mat, d, pos, normal, rng, att, outDir);
                            if (!ok) { col = col + thr * mat.emissive; finished = true; }
                            else
                            {
                                if (doMatE) col = col + thr * mat.emissive;
                                thr = thr * att;
                                const int depth = ((m2 >> 8) & 0xff) + 1;
                                m2 = 0 | (depth << 8) | (1 << 16);
                                WFLD(WF_O, slot) = make_float4(pos.x, pos.y, pos.z, 0.0f);
                                fdd = make_float4(outDir.x, outDir.y, outDir.z, 0.0f);
                            }
                        }
                    }
                    if (finished)
                    {
                        const float w = WFLD(WF_NEXT, slot).w;
                        red_add_f4(p.image + (size_t)__float_as_uint(fcol.w) * 4, col.x * w, col.y * w, col.z * w);
                        m2 = kKindFree;
                    }
                    fdd.w = __int_as_float(m2);
                    WFLD(WF_D, slot) = fdd;
                    WFLD(WF_COL, slot) = make_float4(col.x, col.y, col.z, fcol.w);
                    WFLD(WF_THR, slot) = make_float4(thr.x, thr.y, thr.z, __uint_as_float(rng));
                }
                const unsigned lm = __ballot_sync(0xffffffffu, wantLight);
                if (lm)
                {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&sLightCount, __popc(lm));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (wantLight) sLight[base + __popc(lm & ltMask)] = (uint16_t)slot;
                }
                chunk += kWaveThreads / 32;
            }
        }
        __syncthreads();

        // ---- phase 4: explicit light sampling for Lambert vertices / returning shadow rays (Test.cpp:96-133)
        {
            const int nL = sLightCount;
            for (int k = tid; k < ((nL + 31) & ~31); k += kWaveThreads)
            {
                if (k < nL)
                {
                    const int slot = sLight[k];
                    float4 fdd = WFLD(WF_D, slot);
                    int m2 = __float_as_int(fdd.w);
                    const float4 fta = WFLD(WF_THRALB, slot);
                    const int mid = __float_as_int(fta.w);
                    int j = m2 & 0xff;                 // kind k = shadow ray of light k-1 just returned -> next is light k
                    while (j < sc.nLights && sc.lights[j].id == mid) ++j;
                    if (j < sc.nLights)
                    {
                        const LightRec Lr = sc.lights[j];
                        const float4 fo = WFLD(WF_O, slot), fnl = WFLD(WF_NL, slot), fal = WFLD(WF_ALB, slot);
                        float4 fthr = WFLD(WF_THR, slot);
                        uint32_t rng = __float_as_uint(fthr.w);
                        const V3 o = v3(fo.x, fo.y, fo.z);
                        V3 pc = v3(Lr.cx, Lr.cy, Lr.cz) - o;
                        float d2 = dot(pc, pc);
                        float inv = rsqrtf(d2);
                        V3 sw = pc * inv;
                        V3 su = M<false>::normalize(cross(fabsf(sw.x) > 0.01f ? v3(0, 1, 0) : v3(1, 0, 0), sw));
                        V3 sv = cross(sw, su);
                        float cosAMax = M<false>::sqrt_(1.0f - Lr.radius * Lr.radius * inv * inv);
                        float eps1 = RandomFloat01(rng), eps2 = RandomFloat01(rng);
                        float cosA = 1.0f - eps1 + eps1 * cosAMax;
                        float sinA = M<false>::sqrt_(1.0f - cosA * cosA);
                        float sp, cp;
                        __sincosf(2.0f * TPT_PI * eps2, &sp, &cp);
                        V3 l = su * (cp * sinA) + sv * (sp * sinA) + sw * cosA;
                        float omega = 2.0f * TPT_PI * (1.0f - cosAMax);
                        float dl = dot(l, v3(fnl.x, fnl.y, fnl.z));
                        float m = (0.0f < dl) ? dl : 0.0f;
                        V3 pend = v3(fthr.x, fthr.y, fthr.z) * ((v3(fal.x, fal.y, fal.z) * v3(Lr.ex, Lr.ey, Lr.ez)) * (m * omega * (1.0f / TPT_PI)));
                        WFLD(WF_PEND, slot) = make_float4(pend.x, pend.y, pend.z, 0.0f);
                        fthr.w = __uint_as_float(rng);
                        WFLD(WF_THR, slot) = fthr;
                        m2 = (m2 & ~0xff) | (1 + j);
                        WFLD(WF_D, slot) = make_float4(l.x, l.y, l.z, __int_as_float(m2));
                    }
                    else
                    {
                        const float4 fnx = WFLD(WF_NEXT, slot);
                        float4 fthr = WFLD(WF_THR, slot);
                        const int depth = ((m2 >> 8) & 0xff) + 1;
                        m2 = 0 | (depth << 8) | (0 << 16);       // doMaterialE = false after a Lambert vertex (Test.cpp:214)
                        WFLD(WF_D, slot) = make_float4(fnx.x, fnx.y, fnx.z, __int_as_float(m2));
                        WFLD(WF_THR, slot) = make_float4(fta.x, fta.y, fta.z, fthr.w);
                    }
                }
            }
        }
        __syncthreads();
    }
#undef WFLD
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(p.rayCounter, (unsigned long long)rc);
}

} // namespace tpt
#include "tpt_refgpu.cuh"
namespace tpt {

// The reference-GPU-compatible mode's native instance sweeps like the fast mode: expanded form on packed pairs (FFMA2).
struct RgSetupK2
{
    using Hitter = FastHitterK2;
    static constexpr int kMinBlocks = 6;
    static size_t extra_smem(const SceneDev& sc) { return (((size_t)sc.stagedBytes + 127u) & ~(size_t)127u) - sc.stagedBytes + (size_t)((sc.count + 3) / 4 * 4) * 16; }
    static __device__ __forceinline__ Hitter prepare(SceneView& sc, unsigned char* smem, const SceneBlobLayout& L, uint32_t stagedBytes)
    {
        float4* sphK = reinterpret_cast<float4*>(smem + L.offSph);
        build_sphK(sc, sphK);                       // padded spheres get K = 1e30: never candidates (ComputeShader.hlsl:134)
        __syncthreads();
        float4* pairs = reinterpret_cast<float4*>(smem + ((stagedBytes + 127u) & ~127u));
        build_sph_pairs(sc, sphK, pairs);
        __syncthreads();
        Hitter h; h.sphK = sc.sphShared; h.sphP = smem_u32(pairs); h.simdCount = sc.simdCount;
        asm volatile("" : "+r"(h.sphK), "+r"(h.sphP));
        return h;
    }
};

cudaError_t launch_refgpu_fast(const DrawParams& p, const SceneDev& sc, int numSMs, cudaStream_t stream)
{
    const int simd = (sc.count + 3) / 4 * 4;
    if (sc.kformOk && sc.kformMode == 2 && simd <= 1024) return launch_refgpu_t<false, RgSetupK2>(p, sc, numSMs, stream);
    return launch_refgpu_t<false>(p, sc, numSMs, stream);
}

template <int THREADS, int MINB, int KFORM>
static cudaError_t launch_queue_t(const DrawParams& p, const SceneDev& sc, int numSMs, cudaStream_t stream,
                                  unsigned int* bandDone, int numBands, unsigned int* bandExpected)
{
    auto kern = k_fast_queue<THREADS, MINB, KFORM>;
    const int simd = (sc.count + 3) / 4 * 4;
    const size_t stagedAl = ((size_t)sc.stagedBytes + 127u) & ~(size_t)127u;
    const size_t raysOffset = stagedAl + (KFORM >= 2 ? (((size_t)simd * 16 + 127u) & ~(size_t)127u) : 0);
    const size_t dyn3 = raysOffset + (size_t)(THREADS / 32) * kSlabPix * 32;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn3);
    if (e != cudaSuccess) return e;
    int perSM = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, THREADS, dyn3);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    float wPrev = 1.0f;
    for (int f = 0; f < p.numFrames; ++f) wPrev *= lerp_fac(p.frame0 + f, p.flags);
    const long long regionPix = (long long)p.numRows * p.width;
    k_prepare_image<<<(unsigned)((regionPix + 255) / 256), 256, 0, stream>>>(p, wPrev);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const uint32_t S = (uint32_t)(p.spp * p.numFrames);
    const long long slabs = ((regionPix + kSlabPix - 1) / kSlabPix) * S;
    if (slabs > 0x7fffffffLL) return cudaErrorInvalidValue;
    long long grid = (long long)numSMs * perSM;
    const long long ctasNeeded = (slabs + THREADS / 32 - 1) / (THREADS / 32);
    if (grid > ctasNeeded) grid = ctasNeeded;
    e = cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    // optional progress bands (host-buffer draws): band b = macro-tiles [b*mpb, (b+1)*mpb)
    const uint32_t mtiles = (uint32_t)((regionPix + kSlabPix - 1) / kSlabPix);
    uint32_t mpb = 0;
    if (bandDone && numBands > 0)
    {
        mpb = (mtiles + (uint32_t)numBands - 1) / (uint32_t)numBands;
        for (int b = 0; b < numBands; ++b)
        {
            const long long p0 = (long long)b * mpb * kSlabPix, p1 = (long long)(b + 1) * mpb * kSlabPix;
            const long long px = (p1 < regionPix ? p1 : regionPix) - (p0 < regionPix ? p0 : regionPix);
            bandExpected[b] = (unsigned int)(px * S);
        }
    }
#if TPT_TRACE_WARPS
    static unsigned long long* dTrace = nullptr;
    const size_t traceWords = (size_t)grid * (THREADS / 32) * 8;
    if (getenv("TPT_TRACE_FILE"))
    {
        if (!dTrace) { cudaMalloc(&dTrace, (size_t)148 * 16 * 32 * 8 * 8); cudaMemcpyToSymbol(g_warpTrace, &dTrace, sizeof(dTrace)); }
        cudaMemsetAsync(dTrace, 0, traceWords * 8, stream);
    }
#endif
    kern<<<(unsigned)grid, THREADS, dyn3, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes,
                                                      (uint32_t)slabs, S, mpb ? bandDone : nullptr, mpb ? mpb : 1u, (uint32_t)raysOffset);
#if TPT_TRACE_WARPS
    if (const char* tf = getenv("TPT_TRACE_FILE"))
    {
        cudaStreamSynchronize(stream);
        std::vector<unsigned long long> h(traceWords);
        cudaMemcpy(h.data(), dTrace, traceWords * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(tf, "wb")) { fwrite(h.data(), 8, traceWords, f); fclose(f); }     // the last traced launch wins
    }
#endif
    return cudaGetLastError();
}

// Diagnostic (tpt_debug_hit): the nearest-hit query of one sweep form on caller-supplied rays, staged exactly like
// k_fast_queue stages it. The parity tests compare forms ray by ray with it (ids and distance bits).
template <int KFORM>
__global__ void __launch_bounds__(256)
k_debug_hit(const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights, uint32_t stagedBytes,
            const float* __restrict__ rays, int* __restrict__ outId, float* __restrict__ outT, long long n)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    stage_blob(smem, blob, stagedBytes, &bar);
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);
    float4* sphK = reinterpret_cast<float4*>(smem + L.offSph);
    if (KFORM == 1 || KFORM == 2) build_sphK(sc, sphK);
    __syncthreads();
    float4* pairs = reinterpret_cast<float4*>(smem + ((stagedBytes + 127u) & ~127u));
    if (KFORM == 2) { build_sph_pairs(sc, sphK, pairs); __syncthreads(); }
    if (KFORM == 3) { build_sph_pairs_from_r2(sc, sphK, pairs); __syncthreads(); }
    FastHitterK hitK; hitK.sphK = sc.sphShared; hitK.simdCount = sc.simdCount;
    FastHitterK2 hitK2; hitK2.sphK = sc.sphShared; hitK2.sphP = smem_u32(pairs); hitK2.simdCount = sc.simdCount;
    FastHitterK2C hitK2C; hitK2C.sph = sc.sphShared; hitK2C.sphP = smem_u32(pairs); hitK2C.simdCount = sc.simdCount;
    asm volatile("" : "+r"(hitK.sphK), "+r"(hitK2.sphK), "+r"(hitK2.sphP), "+r"(hitK2C.sph), "+r"(hitK2C.sphP));
    SerialHitter<false> hitS;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
        const V3 o{rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]}, d{rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]};
        float t = TPT_MAX_T;
        int id;
        if (KFORM == 3) id = hitK2C.hit(sc, o, d, TPT_MIN_T, TPT_MAX_T, t);
        else if (KFORM == 2) id = hitK2.hit(sc, o, d, TPT_MIN_T, TPT_MAX_T, t);
        else if (KFORM == 1) id = hitK.hit(sc, o, d, TPT_MIN_T, TPT_MAX_T, t);
        else id = hitS.hit(sc, o, d, TPT_MIN_T, TPT_MAX_T, t);
        outId[i] = id;
        outT[i] = t;
    }
}

cudaError_t launch_debug_hit(const SceneDev& sc, int kform, const float* dRays, int* dId, float* dT, long long n, int numSMs, cudaStream_t stream)
{
    auto kern = kform == 3 ? k_debug_hit<3> : kform == 2 ? k_debug_hit<2> : kform == 1 ? k_debug_hit<1> : k_debug_hit<0>;
    const int simd = (sc.count + 3) / 4 * 4;
    const size_t stagedAl = ((size_t)sc.stagedBytes + 127u) & ~(size_t)127u;
    const size_t dyn = stagedAl + (kform >= 2 ? (size_t)simd * 16 : 0);
    if (dyn > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    kern<<<numSMs, 256, dyn, stream>>>(sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes, dRays, dId, dT, n);
    return cudaGetLastError();
}

// Sweep form the slab-queue kernel (variants 3/4) runs for this scene and option set: see k_fast_queue.
int fast_queue_kform(const SceneDev& sc)
{
    const int simd = (sc.count + 3) / 4 * 4;
    const size_t pairBytes = (size_t)simd * 16, stagedAl = ((size_t)sc.stagedBytes + 127u) & ~(size_t)127u;
    const bool pairsFit = stagedAl + pairBytes + 24 * 2048 <= 200 * 1024;
    if (sc.kformMode == 0) return 0;
    if (sc.kformOk) return (sc.kformMode == 1 || !pairsFit) ? 1 : 2;
    return (sc.kformMode == 2 && pairsFit) ? 3 : 0;
}

int fast_slab_pixels() { return kSlabPix; }

int fast_kernel_launches(const DrawParams&, int variant) { return (variant == 3 || variant == 4 || variant == 9 || variant == 6 || variant == 7) ? 2 : 1; }
bool fast_variant_writes_final_pixels(int variant) { return variant == 8; }



cudaError_t launch_fast(const DrawParams& p, const SceneDev& sc, int variant, int numSMs, cudaStream_t stream,
                        unsigned int* bandDone, int numBands, unsigned int* bandExpected)
{
    if (p.numFrames > kMaxFramesPerDraw) return cudaErrorInvalidValue;
    cudaError_t e;
    if (variant == 0)
    {
        e = cudaFuncSetAttribute(k_fast_mega, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.stagedBytes);
        if (e != cudaSuccess) return e;
        const int tilesX = (p.width + 15) / 16, tilesY = (p.numRows + 15) / 16;
        k_fast_mega<<<tilesX * tilesY, kFastThreads, sc.stagedBytes, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes);
        return cudaGetLastError();
    }
    if (variant == 1 || variant == 2)
    {
        auto kern = variant == 1 ? k_fast_persistent<2> : k_fast_persistent<3>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.stagedBytes);
        if (e != cudaSuccess) return e;
        int perSM = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kFastThreads, sc.stagedBytes);
        if (e != cudaSuccess) return e;
        if (perSM < 1) perSM = 1;
        const long long regionPix = (long long)p.numRows * p.width;
        const int numTiles = (int)((regionPix + kTilePix - 1) / kTilePix);
        int grid = numSMs * perSM;
        if (grid > numTiles) grid = numTiles;
        e = cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess) return e;
        kern<<<grid, kFastThreads, sc.stagedBytes, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes, numTiles);
        return cudaGetLastError();
    }
    if (variant == 3 || variant == 4 || variant == 9)
    {
        // sweep form: 2 = expanded + packed pairs (FFMA2; the pair array costs 16 B per sphere of extra shared memory) where
        // the scene passes the accuracy gate, 3 = the conservative packed form + reference-form pass 2 where it does not,
        // 1 / 0 = scalar expanded / reference form (comparison, or scenes too large for the pair array)
        const int simd = (sc.count + 3) / 4 * 4;
        const size_t pairBytes = (size_t)simd * 16, stagedAl = ((size_t)sc.stagedBytes + 127u) & ~(size_t)127u;
        const int kform = fast_queue_kform(sc);
        // small scenes: 128-thread CTAs, 6 (variant 4: 7) per SM; scenes whose arrays fill most of an SM's shared memory: ONE
        // 768-thread CTA per SM (24 warps share one copy of the geometry instead of 2 CTAs x 4 warps with a copy each)
        const bool big = stagedAl + ((kform >= 2) ? pairBytes : 0) + 4 * 2048 > 36 * 1024 || sc.stagedBytes != sc.layout.totalBytes;
#define TPT_LAUNCH_QUEUE(T, M, K) return launch_queue_t<T, M, K>(p, sc, numSMs, stream, bandDone, numBands, bandExpected)
        if (big) { switch (kform) { case 3: TPT_LAUNCH_QUEUE(TPT_BIG_THREADS, 1, 3); case 2: TPT_LAUNCH_QUEUE(TPT_BIG_THREADS, 1, 2); case 1: TPT_LAUNCH_QUEUE(TPT_BIG_THREADS, 1, 1); default: TPT_LAUNCH_QUEUE(TPT_BIG_THREADS, 1, 0); } }
        // long draws (many slabs per warp) run 8 CTAs of 64 registers per SM instead of 6 of 80: measured 26.2 vs 25.3 Gray/s at
        // 3840x2160x16spp, but 21.6 vs 21.9 at 1280x720x4spp, where a warp only gets ~12 slabs and the drain of the last ones
        // weighs more with more warps. Variant 9 forces the 8-CTA instance (A/B runs).
        const long long slabsPerWarp = (((long long)p.numRows * p.width + kSlabPix - 1) / kSlabPix) * p.spp * p.numFrames / ((long long)numSMs * 32);
        if (kform == 2 && (variant == 9 || (variant == 3 && slabsPerWarp >= 48))) TPT_LAUNCH_QUEUE(128, 8, 2);
        if (variant == 3 || variant == 9) { switch (kform) { case 3: TPT_LAUNCH_QUEUE(128, TPT_QUEUE_MINB, 3); case 2: TPT_LAUNCH_QUEUE(128, TPT_QUEUE_MINB, 2); case 1: TPT_LAUNCH_QUEUE(128, TPT_QUEUE_MINB, 1); default: TPT_LAUNCH_QUEUE(128, TPT_QUEUE_MINB, 0); } }
        switch (kform) { case 3: TPT_LAUNCH_QUEUE(128, 7, 3); case 2: TPT_LAUNCH_QUEUE(128, 7, 2); case 1: TPT_LAUNCH_QUEUE(128, 7, 1); default: TPT_LAUNCH_QUEUE(128, 7, 0); }   // variant 4: 72 registers, 21.05 vs 21.14 Gray/s
#undef TPT_LAUNCH_QUEUE
    }
    if (variant == 8)
    {
        const int simd = (sc.count + 3) / 4 * 4;
        const int kform = !sc.kformOk ? 0 : (sc.kformMode == 1 || simd > 1024 ? 1 : 2);
        const bool alls = sc.stagedBytes == sc.layout.totalBytes;
        auto kern = alls ? (kform == 2 ? k_fast_group<TPT_QUEUE_MINB, 2, true> : kform == 1 ? k_fast_group<TPT_QUEUE_MINB, 1, true> : k_fast_group<TPT_QUEUE_MINB, 0, true>)
                         : (kform == 2 ? k_fast_group<TPT_QUEUE_MINB, 2, false> : kform == 1 ? k_fast_group<TPT_QUEUE_MINB, 1, false> : k_fast_group<TPT_QUEUE_MINB, 0, false>);
        const size_t dyn8 = kform == 2 ? ((sc.stagedBytes + 127u) & ~127u) + (size_t)simd * 16 : sc.stagedBytes;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn8);
        if (e != cudaSuccess) return e;
        int perSM = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kQueueThreads, dyn8);
        if (e != cudaSuccess) return e;
        if (perSM < 1) perSM = 1;
        float wPrev = 1.0f;
        for (int f = 0; f < p.numFrames; ++f) wPrev *= lerp_fac(p.frame0 + f, p.flags);
        const long long regionPix = (long long)p.numRows * p.width;
        const uint32_t S = (uint32_t)(p.spp * p.numFrames);
        const long long groups = (regionPix + kGroupPix - 1) / kGroupPix;
        if (groups > 0x7fffffffLL) return cudaErrorInvalidValue;
        long long grid = (long long)numSMs * perSM;
        const long long ctasNeeded = (groups + kQueueThreads / 32 - 1) / (kQueueThreads / 32);
        if (grid > ctasNeeded) grid = ctasNeeded;
        e = cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)grid, kQueueThreads, dyn8, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes,
                                                                (uint32_t)groups, S, wPrev);
        return cudaGetLastError();
    }
    if (variant == 6 || variant == 7)
    {
        auto kern = variant == 6 ? k_fast_wave<3> : k_fast_wave<4>;
        const uint32_t poolOffset = (sc.stagedBytes + 127u) & ~127u;
        const size_t dyn = (size_t)poolOffset + (size_t)kWaveFields * kWaveSlots * 16;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        int perSM = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kWaveThreads, dyn);
        if (e != cudaSuccess) return e;
        if (perSM < 1) perSM = 1;
        float wPrev = 1.0f;
        for (int f = 0; f < p.numFrames; ++f) wPrev *= lerp_fac(p.frame0 + f, p.flags);
        const long long regionPix = (long long)p.numRows * p.width;
        k_prepare_image<<<(unsigned)((regionPix + 255) / 256), 256, 0, stream>>>(p, wPrev);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        WaveArgs wa;
        wa.S = (uint32_t)(p.spp * p.numFrames);
        const long long total = ((regionPix + 127) / 128) * 128 * wa.S;      // whole 128-pixel slabs; tail pixels are skipped
        if (total > 0x7fffffffLL) return cudaErrorInvalidValue;
        wa.totalPaths = (uint32_t)total;
        wa.divS = make_fastdiv(wa.S); wa.divW = make_fastdiv((uint32_t)p.width); wa.divSpp = make_fastdiv((uint32_t)p.spp);
        long long grid = (long long)numSMs * perSM;
        const long long need = (total + kWaveChunk - 1) / kWaveChunk;
        if (grid > need) grid = need;
        e = cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)grid, kWaveThreads, dyn, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes, poolOffset, wa);
        return cudaGetLastError();
    }
    if (variant == 5)
    {
        const int simd = (sc.count + 3) / 4 * 4;
        const int kform = !sc.kformOk ? 0 : (sc.kformMode == 1 || simd > 1024 ? 1 : 2);
        auto kern = kform == 2 ? k_fast_tileq<6, 2> : kform == 1 ? k_fast_tileq<6, 1> : k_fast_tileq<6, 0>;
        const size_t dyn5 = kform == 2 ? ((sc.stagedBytes + 127u) & ~127u) + (size_t)simd * 16 : sc.stagedBytes;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn5);
        if (e != cudaSuccess) return e;
        int perSM = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kQueueThreads, dyn5);
        if (e != cudaSuccess) return e;
        if (perSM < 1) perSM = 1;
        float wPrev = 1.0f;
        for (int f = 0; f < p.numFrames; ++f) wPrev *= lerp_fac(p.frame0 + f, p.flags);
        const long long regionPix = (long long)p.numRows * p.width;
        // tile size: >= ~8192 paths per tile keeps the per-tile barrier tail < 1-2 %; then shrink (down to one slab)
        // until there are >= 8 tiles per CTA so the end-of-kernel quantisation stays small on sharded images
        const uint32_t S = (uint32_t)(p.spp * p.numFrames);
        long long grid = (long long)numSMs * perSM;
        uint32_t tileQPix = (8192 / S + kSlabPix - 1) / kSlabPix * kSlabPix;
        if (tileQPix < (uint32_t)kSlabPix) tileQPix = kSlabPix;
        if (tileQPix > (uint32_t)kTileQPix) tileQPix = kTileQPix;
        while (tileQPix > (uint32_t)kSlabPix && (regionPix + tileQPix - 1) / tileQPix < 8 * grid && tileQPix * S > 2048) tileQPix -= kSlabPix;
        const long long tiles = (regionPix + tileQPix - 1) / tileQPix;
        if (grid > tiles) grid = tiles;
        e = cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)grid, kQueueThreads, dyn5, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes,
                                                                (uint32_t)tiles, S, wPrev, tileQPix);
        return cudaGetLastError();
    }
    return cudaErrorInvalidValue;
}

} // namespace tpt
