// Internal interface between the C-ABI (tpt_api.cu) and the two kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "tpt_types.h"

namespace tpt {

struct SceneDev
{
    const unsigned char* blob;   // device copy of the packed scene (tpt_scene_pack.h)
    SceneBlobLayout layout;
    int count, nLights;
    uint32_t stagedBytes;        // prefix of the blob every CTA stages into shared memory via TMA
    int kformMode;               // fast kernels: 2 = expanded form with packed pairs (FFMA2) where it fits, 1 = expanded form, scalar
    bool kformOk;                // every sphere satisfies |c|^2 <= 128 + 2 r^2: the expanded-form sweep of the fast kernels
                                 // then rounds no worse than the reference form (see FastHitterK)
};

// lanes: 0 = choose from the number of (frame,row) chains; 1, 8 or 32 to force.
// resolve = false: with numFrames > 1 only the per-frame colours are written to p.scratch (no blend into the image)
cudaError_t launch_exact(const DrawParams& p, const SceneDev& sc, int lanes, cudaStream_t stream, bool resolve = true);
// the reference's progressive blend (Test.cpp:272-276,293-295) of p.numFrames frames of per-frame colours in p.scratch
cudaError_t launch_resolve_exact(const DrawParams& p, cudaStream_t stream);
cudaError_t launch_debug_libm(int fn, const float* dIn, float* dOut, long long n, cudaStream_t stream);
// sweep form variants 3/4 run for this scene (0 reference, 1 expanded, 2 packed pairs, 3 conservative packed)
int fast_queue_kform(const SceneDev& sc);
// nearest hit of n rays {o.xyz, d.xyz} with one sweep form of the fast kernels (0 reference, 1 expanded, 2 packed pairs, 3 conservative)
cudaError_t launch_debug_hit(const SceneDev& sc, int kform, const float* dRays, int* dId, float* dT, long long n, int numSMs, cudaStream_t stream);
// variant: see tpt_fast.cu
// bandDone (optional, variant 3/4 only): device counters [numBands]; the kernel publishes finished paths per band of
// macro-tiles, bandExpected[b] receives the final count of band b (host array) and kSlabPix*ceil(mtiles/numBands) pixels
// form a band.
cudaError_t launch_fast(const DrawParams& p, const SceneDev& sc, int variant, int numSMs, cudaStream_t stream,
                        unsigned int* bandDone = nullptr, int numBands = 0, unsigned int* bandExpected = nullptr);
// TPT_MODE_REFGPU (tpt_refgpu.cuh): strict arithmetic (tpt_exact.cu) / GPU-native arithmetic (tpt_fast.cu)
cudaError_t launch_refgpu_exact(const DrawParams& p, const SceneDev& sc, int numSMs, cudaStream_t stream);
cudaError_t launch_refgpu_fast(const DrawParams& p, const SceneDev& sc, int numSMs, cudaStream_t stream);
int fast_slab_pixels();
int fast_kernel_launches(const DrawParams& p, int variant);
bool fast_variant_writes_final_pixels(int variant);   // the trace kernel stores finished pixels itself (no L2 reductions)

} // namespace tpt
