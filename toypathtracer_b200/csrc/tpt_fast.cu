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
#include "tpt_integrator.cuh"
#include "tpt_device_utils.cuh"
#include "tpt_launch.h"

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
                Ray r = GetRay<false>(p.cam, u, v, state);
                col = col + trace_fast(sc, r, state, rc, hitter);
            }
            acc = acc + col * (invSpp * sW[fi]);
        }
        float* px = p.image + ((size_t)(p.packed ? ri : y) * p.width + x) * 4;
        float4 prev = make_float4(0, 0, 0, 0);
        const float wPrev = sWPrev;
        if (wPrev != 0.0f) prev = ld_stream_f4(px);
        prev.x = prev.x * wPrev + acc.x; prev.y = prev.y * wPrev + acc.y; prev.z = prev.z * wPrev + acc.z;
        st_stream_f4(px, prev);
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
                        Ray r = GetRay<false>(p.cam, u, v, st.rng);
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
                    if (id < 0) { st.col = st.col + st.thr * sky(st.d); finished = true; }
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
            float4 prev = make_float4(0, 0, 0, 0);
            if (wPrev != 0.0f) prev = ld_stream_f4(px);
            prev.x = prev.x * wPrev + sAcc[i * 3 + 0];
            prev.y = prev.y * wPrev + sAcc[i * 3 + 1];
            prev.z = prev.z * wPrev + sAcc[i * 3 + 2];
            st_stream_f4(px, prev);
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
constexpr int kSlabPix = 128;
constexpr int kQueueThreads = 128;

__global__ void k_prepare_image(DrawParams p, float wPrev)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)p.numRows * p.width;
    if (idx >= n) return;
    const int ri = (int)(idx / p.width), x = (int)(idx % p.width);
    const int y = p.row0 + ri * p.rowStep;
    float* px = p.image + ((size_t)(p.packed ? ri : y) * p.width + x) * 4;
    float4 v = make_float4(0, 0, 0, 0);
    if (wPrev != 0.0f)
    {
        v = ld_stream_f4(px);
        v.x *= wPrev; v.y *= wPrev; v.z *= wPrev;
    }
    *reinterpret_cast<float4*>(px) = v;   // stays in L2 for the reductions that follow
}

__device__ __forceinline__ void red_add_f4(float* addr, float x, float y, float z)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(0.0f) : "memory");
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

// One iteration of the per-lane path state machine shared by the queue kernels: intersect the lane's current ray
// (path or shadow) against all spheres, then shade. Returns true when the lane's path has ended (st.col is final).
__device__ __forceinline__ bool path_step(const SceneView& sc, QPath& st, unsigned& rc)
{
    SerialHitter<false> hitter;
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
            if (id < 0) { st.col = st.col + st.thr * sky(st.d); finished = true; }
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
                else if (mat.type == kMetal)
                {
                    // Test.cpp:137-150; with roughness == 0 the unit-sphere sample has zero weight, so the fast
                    // mode skips drawing it (the exact mode must draw it: it advances the shared RNG stream)
                    V3 refl = reflect(st.d, normal);
                    if (mat.roughness != 0.0f) refl = refl + mat.roughness * RandomInUnitSphere(st.rng);
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
            float inv = rsqrtf(d2);
            V3 sw = pc * inv;
            V3 su = M<false>::normalize(cross(fabsf(sw.x) > 0.01f ? v3(0, 1, 0) : v3(1, 0, 0), sw));
            V3 sv = cross(sw, su);
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

template <int MINB>
__global__ void __launch_bounds__(kQueueThreads, MINB)
k_fast_queue(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights,
             uint32_t stagedBytes, uint32_t numSlabs, uint32_t S)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ float sW[kMaxFramesPerDraw];
    stage_blob(smem, blob, stagedBytes, &bar);
    if (threadIdx.x == 0) { float wp; blend_weights(p, sW, wp); }
    __syncthreads();
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);
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
                const uint32_t mtile = slab / S, s = slab - mtile * S;
                const uint32_t pix0 = mtile * kSlabPix;
                slabEnd = regionPix - pix0 < (uint32_t)kSlabPix ? regionPix - pix0 : (uint32_t)kSlabPix;
                slabCur = 0;
                slabRi0 = (int)(pix0 / (uint32_t)p.width);
                slabX0 = (int)(pix0 - (uint32_t)slabRi0 * (uint32_t)p.width);
                const uint32_t fi = s / (uint32_t)p.spp;
                slabSample = s - fi * (uint32_t)p.spp;
                slabFrame = (uint32_t)p.frame0 + fi;
                slabW = invSpp * sW[fi];
            }
            const uint32_t avail = slabEnd - slabCur;
            const uint32_t rank = (uint32_t)__popc(need & ltMask);
            if (!st.active && rank < avail)
            {
                int x = slabX0 + (int)(slabCur + rank), ri = slabRi0;
                while (x >= p.width) { x -= p.width; ++ri; }
                const int y = p.row0 + ri * p.rowStep;
                st.rng = pixel_seed((uint32_t)(y * p.width + x) * (uint32_t)p.spp + slabSample, slabFrame);
                float u = ((float)x + RandomFloat01(st.rng)) * p.invWidth;
                float v = ((float)y + RandomFloat01(st.rng)) * p.invHeight;
                Ray r = GetRay<false>(p.cam, u, v, st.rng);
                st.o = r.orig; st.d = r.dir;
                st.thr = v3(1, 1, 1); st.col = v3(0, 0, 0);
                st.pixOff = (uint32_t)((p.packed ? ri : y) * p.width + x);
                st.weight = slabW;
                st.kind = 0; st.depth = 0; st.doMaterialE = true; st.active = true;
            }
            const uint32_t n = (uint32_t)__popc(need);
            slabCur += n < avail ? n : avail;
            need = __ballot_sync(0xffffffffu, !st.active);
        }
        if (!__any_sync(0xffffffffu, st.active)) break;

        const bool finished = path_step(sc, st, rc);
        if (finished)
        {
            red_add_f4(p.image + (size_t)st.pixOff * 4, st.col.x * st.weight, st.col.y * st.weight, st.col.z * st.weight);
            st.active = false;
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
constexpr int kTileQPix = 2048;

template <int MINB>
__global__ void __launch_bounds__(kQueueThreads, MINB)
k_fast_tileq(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights,
             uint32_t stagedBytes, uint32_t numTiles, uint32_t S, float wPrev, uint32_t tileQPix)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ float sW[kMaxFramesPerDraw];
    __shared__ float sAcc[kTileQPix * 3];
    __shared__ uint32_t sTile, sSlab;
    stage_blob(smem, blob, stagedBytes, &bar);
    if (threadIdx.x == 0) { float wp; blend_weights(p, sW, wp); }
    for (int i = threadIdx.x; i < kTileQPix * 3; i += kQueueThreads) sAcc[i] = 0.0f;
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);
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
                }
                const uint32_t avail = slabEnd - slabCur;
                const uint32_t rank = (uint32_t)__popc(need & ltMask);
                if (!st.active && rank < avail)
                {
                    const uint32_t q = slabCur + rank;
                    int x = slabX0 + (int)q, ri = slabRi0;
                    while (x >= p.width) { x -= p.width; ++ri; }
                    const int y = p.row0 + ri * p.rowStep;
                    st.rng = pixel_seed((uint32_t)(y * p.width + x) * (uint32_t)p.spp + slabSample, slabFrame);
                    float u = ((float)x + RandomFloat01(st.rng)) * p.invWidth;
                    float v = ((float)y + RandomFloat01(st.rng)) * p.invHeight;
                    Ray r = GetRay<false>(p.cam, u, v, st.rng);
                    st.o = r.orig; st.d = r.dir;
                    st.thr = v3(1, 1, 1); st.col = v3(0, 0, 0);
                    st.pixOff = slabQ0 + q;                 // pixel index inside the tile
                    st.weight = slabW;
                    st.kind = 0; st.depth = 0; st.doMaterialE = true; st.active = true;
                }
                const uint32_t n = (uint32_t)__popc(need);
                slabCur += n < avail ? n : avail;
                need = __ballot_sync(0xffffffffu, !st.active);
            }
            if (!__any_sync(0xffffffffu, st.active)) break;
            if (path_step(sc, st, rc))
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
            float4 prev = make_float4(0, 0, 0, 0);
            if (wPrev != 0.0f) prev = ld_stream_f4(px);
            prev.x = prev.x * wPrev + sAcc[i * 3 + 0];
            prev.y = prev.y * wPrev + sAcc[i * 3 + 1];
            prev.z = prev.z * wPrev + sAcc[i * 3 + 2];
            st_stream_f4(px, prev);
            sAcc[i * 3 + 0] = 0.0f; sAcc[i * 3 + 1] = 0.0f; sAcc[i * 3 + 2] = 0.0f;
        }
    }
    for (int off = 16; off > 0; off >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, off);
    if (lane == 0 && rc) atomicAdd(p.rayCounter, (unsigned long long)rc);
}

int fast_kernel_launches(const DrawParams&, int variant) { return (variant == 3 || variant == 4) ? 2 : 1; }


cudaError_t launch_fast(const DrawParams& p, const SceneDev& sc, int variant, int numSMs, cudaStream_t stream)
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
    if (variant == 3 || variant == 4)
    {
        auto kern = variant == 3 ? k_fast_queue<6> : k_fast_queue<8>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.stagedBytes);
        if (e != cudaSuccess) return e;
        int perSM = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kQueueThreads, sc.stagedBytes);
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
        const long long warpsNeeded = (slabs + 3) / 4;   // 4 warps per CTA
        if (grid > warpsNeeded) grid = warpsNeeded;
        e = cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)grid, kQueueThreads, sc.stagedBytes, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes,
                                                                      (uint32_t)slabs, S);
        return cudaGetLastError();
    }
    if (variant == 5)
    {
        auto kern = k_fast_tileq<6>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.stagedBytes);
        if (e != cudaSuccess) return e;
        int perSM = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kQueueThreads, sc.stagedBytes);
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
        kern<<<(unsigned)grid, kQueueThreads, sc.stagedBytes, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes,
                                                                      (uint32_t)tiles, S, wPrev, tileQPix);
        return cudaGetLastError();
    }
    return cudaErrorInvalidValue;
}

} // namespace tpt
