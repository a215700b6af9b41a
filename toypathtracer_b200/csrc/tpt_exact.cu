// Exact ("replay") mode kernels. THIS TRANSLATION UNIT MUST BE COMPILED WITH -fmad=false: the integrator is
// written with plain float operators and relies on the compiler never contracting a*b+c (one
// contraction forks the row's RNG stream). Division and sqrt use the IEEE intrinsics explicitly.
//
// Work decomposition: the reference seeds one XorShift32 stream per (frame,row) and carries it serially across
// the row (Cpp/Source/Test.cpp:278-297), so the independent unit is a CHAIN = (frame, row). LANES lanes of a
// warp cooperate on one chain: all of them carry the (identical) integrator state, the ray-vs-all-spheres
// sweep is split across the lanes (sphere i -> lane i % LANES, SoA from shared memory) and the nearest hit is
// shuffle-reduced. LANES = 32 when chains are scarce (one 4-spp frame = `height` chains), 1 when thousands
// of frames are batched (1024 spp at 720p = 184 320 chains).
#include "tpt_integrator.cuh"
#include "tpt_device_utils.cuh"
#include "tpt_launch.h"

namespace tpt {

template <int LANES>
__global__ void k_trace_exact(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L,
                              int count, int nLights, uint32_t stagedBytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    stage_blob(smem, blob, stagedBytes, &bar);
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);

    const int lane = threadIdx.x & 31;
    const int sub = threadIdx.x % LANES;
    const long long chain = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const long long totalChains = (long long)p.numRows * p.numFrames;
    if (chain >= totalChains) return;
    // consecutive chains = the same row in consecutive frames: neighbouring lanes trace the same pixels with different
    // RNG streams (coherent geometry, similar path lengths), and the expensive bottom rows are scheduled first
    const int ri = (int)(chain / p.numFrames);
    const int fi = (int)(chain % p.numFrames);
    const int y = p.row0 + ri * p.rowStep;
    const int frame = p.frame0 + fi;

    GroupHitter<true, LANES> hitter;
    hitter.sub = sub;
    hitter.mask = LANES == 32 ? 0xffffffffu : (((1u << (LANES & 31)) - 1u) << (lane - sub));

    uint32_t state = row_seed(y, frame);
    unsigned rc = 0;
    const float lerpFac = lerp_fac(frame, p.flags);
    const float oneMinus = 1.0f - lerpFac;
    const size_t imgRow = (size_t)(p.packed ? ri : y) * p.width;
    for (int x = 0; x < p.width; ++x)
    {
        V3 col = pixel_exact(sc, p.cam, x, y, p.spp, p.invWidth, p.invHeight, state, rc, hitter);
        if (sub == 0)
        {
            if (p.numFrames == 1)
            {
                // Test.cpp:293-295: prev*lerpFac + col*(1-lerpFac); alpha is never written by the CPU path
                float4* px = reinterpret_cast<float4*>(p.image + (imgRow + x) * 4);
                float4 prev = *px;
                prev.x = prev.x * lerpFac + col.x * oneMinus;
                prev.y = prev.y * lerpFac + col.y * oneMinus;
                prev.z = prev.z * lerpFac + col.z * oneMinus;
                *px = prev;
            }
            else
            {
                float4* px = reinterpret_cast<float4*>(p.scratch) + ((size_t)fi * p.numRows + ri) * p.width + x;
                *px = make_float4(col.x, col.y, col.z, 0.0f);
            }
        }
    }
    if (sub == 0) atomicAdd(p.rayCounter + fi, (unsigned long long)rc);
}

// One thread per chain, flat form: every loop iteration is one xchain_step() = one sphere sweep, so the lanes of a
// warp (the same row in 32 consecutive frames) stay converged on the sweep whatever their paths are doing.
template <int LANES>
__global__ void __launch_bounds__(128)
k_trace_exact_flat(DrawParams p, const unsigned char* __restrict__ blob, SceneBlobLayout L, int count, int nLights, uint32_t stagedBytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    stage_blob(smem, blob, stagedBytes, &bar);
    SceneView sc = make_view(smem, blob, L, stagedBytes, count, nLights);
    const int lane = threadIdx.x & 31;
    const int sub = threadIdx.x % LANES;
    const long long chain = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const long long totalChains = (long long)p.numRows * p.numFrames;
    if (chain >= totalChains) return;
    const int ri = (int)(chain / p.numFrames);
    const int fi = (int)(chain % p.numFrames);
    const int y = p.row0 + ri * p.rowStep;
    const int frame = p.frame0 + fi;
    const float lerpFac = lerp_fac(frame, p.flags);
    const float oneMinus = 1.0f - lerpFac;
    const size_t imgRow = (size_t)(p.packed ? ri : y) * p.width;
    GroupHitter<true, LANES> hitter;       // LANES lanes share the chain: replicated state, split sweep (see k_trace_exact)
    hitter.sub = sub;
    hitter.mask = LANES == 32 ? 0xffffffffu : (((1u << (LANES & 31)) - 1u) << (lane - sub));
    unsigned rc = 0;
    XChain c;
    xchain_begin(c, p.cam, y, frame, p.invWidth, p.invHeight);
    while (c.x < p.width)
    {
        const int x = c.x;
        V3 col;
        if (xchain_step(sc, p.cam, c, y, p.spp, p.width, p.invWidth, p.invHeight, rc, hitter, col) && sub == 0)
        {
            if (p.numFrames == 1)
            {
                float4* px = reinterpret_cast<float4*>(p.image + (imgRow + x) * 4);
                float4 prev = *px;
                prev.x = prev.x * lerpFac + col.x * oneMinus;
                prev.y = prev.y * lerpFac + col.y * oneMinus;
                prev.z = prev.z * lerpFac + col.z * oneMinus;
                *px = prev;
            }
            else
            {
                float4* px = reinterpret_cast<float4*>(p.scratch) + ((size_t)fi * p.numRows + ri) * p.width + x;
                *px = make_float4(col.x, col.y, col.z, 0.0f);
            }
        }
    }
    if (sub == 0) atomicAdd(p.rayCounter + fi, (unsigned long long)rc);
}

template <int LANES>
static cudaError_t launch_exact_flat_t(const DrawParams& p, const SceneDev& sc, cudaStream_t stream)
{
    const long long threads = (long long)p.numRows * p.numFrames * LANES;
    auto kern = k_trace_exact_flat<LANES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.stagedBytes);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)((threads + 127) / 128), 128, sc.stagedBytes, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes);
    return cudaGetLastError();
}

// Sequential progressive blend of the per-frame colours (Test.cpp:272-276,293-295), one thread per pixel,
// frames in order so the float sequence is the reference's.
__global__ void k_resolve_exact(DrawParams p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)p.numRows * p.width;
    if (idx >= n) return;
    const int ri = (int)(idx / p.width), x = (int)(idx % p.width);
    const int y = p.row0 + ri * p.rowStep;
    float4* px = reinterpret_cast<float4*>(p.image) + (size_t)(p.packed ? ri : y) * p.width + x;
    float4 prev = *px;
    const float4* s = reinterpret_cast<const float4*>(p.scratch) + idx;
    for (int fi = 0; fi < p.numFrames; ++fi, s += n)
    {
        const float lerpFac = lerp_fac(p.frame0 + fi, p.flags);
        const float oneMinus = 1.0f - lerpFac;
        float4 col = ld_stream_f4(reinterpret_cast<const float*>(s));
        prev.x = prev.x * lerpFac + col.x * oneMinus;
        prev.y = prev.y * lerpFac + col.y * oneMinus;
        prev.z = prev.z * lerpFac + col.z * oneMinus;
    }
    *px = prev;
}

template <int LANES>
static cudaError_t launch_exact_t(const DrawParams& p, const SceneDev& sc, cudaStream_t stream, int blockThreads)
{
    const long long totalChains = (long long)p.numRows * p.numFrames;
    const long long threads = totalChains * LANES;
    const int grid = (int)((threads + blockThreads - 1) / blockThreads);
    auto kern = k_trace_exact<LANES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc.stagedBytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, blockThreads, sc.stagedBytes, stream>>>(p, sc.blob, sc.layout, sc.count, sc.nLights, sc.stagedBytes);
    return cudaGetLastError();
}

__global__ void k_debug_libm(int fn, const float* __restrict__ in, float* __restrict__ out, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = in[i];
    out[i] = fn == 0 ? M<true>::sin_(x) : (fn == 1 ? M<true>::cos_(x) : M<true>::pow5_(x));
}

cudaError_t launch_debug_libm(int fn, const float* dIn, float* dOut, long long n, cudaStream_t stream)
{
    k_debug_libm<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(fn, dIn, dOut, n);
    return cudaGetLastError();
}

cudaError_t launch_exact(const DrawParams& p, const SceneDev& sc, int lanes, cudaStream_t stream)
{
    const long long totalChains = (long long)p.numRows * p.numFrames;
    if (lanes <= 0)
    {
        // measured on B200 (profiles/r01, with the REDUX reduction): 720 chains -> 32 lanes (448 vs 177 vs 61 Mray/s for
        // 32/8/1), 11 520 chains -> 32 lanes (2.61 vs 2.48 vs 0.98 Gray/s), 184 320 chains -> 1 lane, flat form (5.8 Gray/s)
        lanes = totalChains >= 100000 ? 1 : (totalChains >= 20000 ? 8 : 32);
    }
    cudaError_t e;
    const long long threads = totalChains * lanes;
    const int block = threads >= 148LL * 256 ? 128 : (threads >= 148LL * 64 ? 64 : 32);
    switch (lanes)
    {
    case 1: e = launch_exact_flat_t<1>(p, sc, stream); break;
    case 2: e = launch_exact_t<1>(p, sc, stream, block); break;      // nested-loop form, one lane per chain (kept for comparison)
    case 8: e = launch_exact_t<8>(p, sc, stream, block); break;
    case 9: e = launch_exact_flat_t<8>(p, sc, stream); break;        // flat form with 8 lanes per chain: measured 2x SLOWER than
                                                                     // the nested form at 11 520 chains (221 vs 111 ms), comparison only
    case 32: e = launch_exact_t<32>(p, sc, stream, block); break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    if (p.numFrames > 1)
    {
        const long long n = (long long)p.numRows * p.width;
        k_resolve_exact<<<(int)((n + 255) / 256), 256, 0, stream>>>(p);
        e = cudaGetLastError();
    }
    return e;
}

} // namespace tpt
