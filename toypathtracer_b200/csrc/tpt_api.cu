// C-ABI of include/tpt_b200.h: context, scene upload, draw dispatch, host<->device plumbing.
// No torch types, no CPU rendering path: every draw ends in a kernel launch or an error.
#include "../../include/tpt_b200.h"
#include "tpt_launch.h"
#include "tpt_scene_pack.h"
#include "tpt_device_utils.cuh"
#include "tpt_integrator.cuh"
#include <cuda.h>      // types only: cuStreamWaitValue32 is resolved at run time (cudaGetDriverEntryPoint), libcuda is not linked
#include <string>
#include <vector>
#include <string.h>
#include <stdio.h>

using namespace tpt;

struct tpt_context
{
    int device = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    bool haveTiming = false;
    std::string lastError;

    // scene: two device blobs + two pinned staging buffers. tpt_set_scene never blocks the host: the new blob goes
    // into the slot the draws in flight are NOT reading, on its own stream (the per-frame scene update of the
    // reference's GPU shells, Cpp/Windows/TestWin.cpp:261-283, as an async double-buffered upload); draws wait on the
    // slot's upload event, uploads wait on the slot's last-use event.
    unsigned char* dBlobs[2] = {nullptr, nullptr};
    size_t blobCap[2] = {0, 0};
    unsigned char* hBlob[2] = {nullptr, nullptr};
    size_t hBlobCap[2] = {0, 0};
    int curBlob = 0;
    cudaStream_t uploadStream = nullptr;
    cudaEvent_t uploadDone[2] = {nullptr, nullptr};
    cudaEvent_t blobLastUse[2] = {nullptr, nullptr};
    std::vector<unsigned char> lastBlob;          // bytes of the blob currently on the device (skip identical uploads)
    cudaEvent_t lastDraw = nullptr;               // end of the most recent draw: draws of one context are serialised in
    cudaStream_t lastDrawStream = nullptr;        // issue order even when they are enqueued on different streams
    bool haveLastDraw = false;
    SceneDev scene{};
    Camera88 cam{};
    bool haveScene = false;
    int spp = 4;

    // options
    int fastVariant = -1;     // -1 = auto: 3 (slab queue + L2 reductions; 7 = material-sorted block wavefront from 1024 spheres) for device buffers, 8 (warp-owned groups, direct coalesced
                              // write-out) when a host-buffer draw can store straight into page-locked memory
    int fastKForm = 2;        // 0: reference-form sweep, 1: expanded form, 2: expanded form with packed pairs (FFMA2); gated per scene by kformOk
    int fastAlphaZero = 0;
    int sceneUploadAlways = 0;   // 1: tpt_set_scene copies the blob even when its bytes are the resident scene's (benchmarks: a real H2D per step)
    long long lastSceneUploadBytes = 0;
    uint32_t sceneFlags = 0;  // kScene* bits for the next tpt_set_scene    // 1: fast-mode draws whose `prev` has zero weight write alpha = 0 instead of preserving it
    int exactLanes = 0;
    // Exact mode, one frame per call (the drop-in's DrawTest): a frame is only `height` serial RNG chains, far too few to
    // fill the GPU (0.76 Gray/s at 720p), while 16 frames at once run at 2.6 Gray/s. With "exact_lookahead" = L > 1 a
    // cache miss traces frames [f, f+L) in ONE launch into a per-frame colour cache and the calls for f+1 .. f+L-1 only
    // blend their cached frame into the caller's buffer (per-frame colours do not depend on the buffer, Test.cpp:283-291):
    // same bits, same per-frame ray counts, L frames of latency on a miss. Not used under kFlagAnimate (the scene of a
    // future frame is not known yet); any scene, camera, size, row-range or spp change invalidates the cache.
    // "exact_lookahead" = -1 is the adaptive form for callers that render frame after frame (the drop-in's DrawTest): the
    // first call traces one frame; each time the caller walks to the end of the cached window and asks for the frame right
    // after it, with nothing else changed, the next window doubles (1, 2, 4, 8, 16 frames). No first-call latency, at most
    // half of the traced frames are speculative when the caller stops or changes anything.
    int exactLookahead = 0;
    int lookLastServed = -1;        // last frame handed out from the current window
    bool lookKeyValid = false;      // `look` describes the last one-frame exact draw (even a window of 1)
    struct LookKey { int frame0, n, width, height, row0, numRows, rowStep, spp; unsigned long long sceneGen; Camera88 cam; } look{};
    bool lookValid = false;
    float* dLook = nullptr; size_t lookCap = 0;
    unsigned long long* dLookRays = nullptr; size_t lookRaysCap = 0;
    unsigned long long sceneGen = 0;
    int registerHost = 0;
    size_t maxScratchBytes = (size_t)8 << 30;

    // buffers
    float* dImage = nullptr; size_t imageCap = 0;      // staging image for host-pointer draws
    float* dScratch = nullptr; size_t scratchCap = 0;  // exact mode per-frame colours
    unsigned long long* dRayCounters = nullptr; int rayCounterCap = 0;   // [numFrames]
    unsigned long long* dAccum = nullptr;              // [0] total since last read, [1] last draw total
    unsigned int* dWork = nullptr;                     // ring of work-counter slots, one per draw (kWorkSlots x 32 uints)
    unsigned workSlot = 0;
    unsigned long long* hPinned = nullptr; int hPinnedCap = 0;
    void* registeredPtr = nullptr; size_t registeredBytes = 0;
    int lastLaunches = 0;

    // host-buffer draws: row bands pipelined over several streams (kernel of band b+1 overlaps D2H of band b)
    static const int kMaxBands = 8;
    static const int kWorkSlots = 32;
    int hostBands = 3;
    cudaStream_t bandStream[kMaxBands] = {};
    cudaEvent_t bandEvent[kMaxBands] = {};
    cudaEvent_t forkEvent = nullptr;

    // progress-triggered D2H (fast variant 3/4): the trace kernel publishes per-band completion counters, the copy
    // stream waits on them with stream memory operations and copies a band while later bands are still being traced
    typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    WaitValue32Fn waitValue32 = nullptr;
    int hostProgress = 1;
    int hostZeroCopy = 1;     // fast variant 8 stores finished pixels straight into page-locked host buffers
    int progressBands = 4;
    unsigned int* dBandDone = nullptr;
    cudaStream_t copyStream = nullptr;
    // diagnostics: timestamps of the last progress-mode draw (kernel end, each band copy end), see tpt_debug_timeline
    cudaEvent_t tlKernelEnd = nullptr, tlBand[16] = {};
    int tlBands = 0;
};

static int fail(tpt_context* ctx, cudaError_t e, const char* what)
{
    if (ctx)
    {
        char buf[512];
        snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
        ctx->lastError = buf;
    }
    return (int)e ? (int)e : -1;
}
static int fail_msg(tpt_context* ctx, const char* msg)
{
    if (ctx) ctx->lastError = msg;
    return (int)cudaErrorInvalidValue;
}
#define CK(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(ctx, _e, what); } while (0)

namespace tpt {
// sums the per-frame counters of one draw into the running totals
// hostMirror (optional): page-locked host memory the caller of tpt_draw reads after the stream sync — the per-frame counts
// arrive there by this kernel's own stores, which saves the separate 8-byte D2H copy (one DMA round trip per draw).
__global__ void k_accumulate_rays(const unsigned long long* perFrame, int n, unsigned long long* accum, unsigned long long* hostMirror)
{
    unsigned long long s = 0;
    for (int i = 0; i < n; ++i) { const unsigned long long v = perFrame[i]; s += v; if (hostMirror) hostMirror[i] = v; }
    accum[0] += s;
    accum[1] = s;
}

// The reference's three presentation conversions of the linear float image to 8 bits per channel (row f.1):
//   transfer 0  LinearToSRGB of the D3D11/Metal presentation pass (Cpp/Windows/PixelShader.hlsl:1-15), rounded
//   transfer 1  min(sqrtf(x) * 255, 255) truncated — the WebAssembly shell's cheap gamma (Cpp/Emscripten/main.cpp:67-79)
//   transfer 2  the C# TGA writer's LinearToSRGB: min((uint)(x * 255.9f), 255) (Cs/Program.cs:61-67)
// flipY: row 0 of the float image is the bottom (the Emscripten shell flips, the TGA writer does not: TGA is bottom-up);
// bgr: TGA channel order (Cs/Program.cs:36-41).
__global__ void k_tonemap_rgba8(const float4* __restrict__ img, int width, int height, uchar4* __restrict__ dst, int transfer, int bgr, int flipY)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= width * height) return;
    const int x = idx % width, y = idx / width;
    float4 c = ld_stream_f4(reinterpret_cast<const float*>(img + (size_t)(flipY ? height - 1 - y : y) * width + x));
    auto enc = [transfer](float v) -> unsigned char {
        if (transfer == 1) return (unsigned char)fminf(__fsqrt_rn(v) * 255.0f, 255.0f);
        v = fmaxf(v, 0.0f);
        v = fmaxf(1.055f * __powf(v, 0.416666667f) - 0.055f, 0.0f);
        if (transfer == 2) { unsigned u = (unsigned)(v * 255.9f); return (unsigned char)(u < 255u ? u : 255u); }
        return (unsigned char)(fminf(v, 1.0f) * 255.0f + 0.5f);
    };
    const unsigned char r = enc(c.x), g = enc(c.y), b = enc(c.z);
    dst[idx] = bgr ? make_uchar4(b, g, r, 255) : make_uchar4(r, g, b, 255);
}
} // namespace tpt

extern "C" {

int tpt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int tpt_create(int device, tpt_context** out)
{
    if (!out) return (int)cudaErrorInvalidValue;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return (int)e;
    if (n <= 0) return (int)cudaErrorNoDevice;   // no CPU fallback, by design
    if (device < 0 || device >= n) return (int)cudaErrorInvalidDevice;
    tpt_context* ctx = new tpt_context();
    ctx->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->evStart);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->evStop);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dAccum, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->dAccum, 0, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dWork, tpt_context::kWorkSlots * 32 * sizeof(unsigned int));   // per draw: 8 bands x 4 uints
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->uploadStream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i)
    {
        e = cudaEventCreateWithFlags(&ctx->uploadDone[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->blobLastUse[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lastDraw, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->forkEvent, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dBandDone, 64);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->tlKernelEnd);
    for (int b = 0; b < 16 && e == cudaSuccess; ++b) e = cudaEventCreate(&ctx->tlBand[b]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            ctx->waitValue32 = (tpt_context::WaitValue32Fn)fn;
        else (void)cudaGetLastError();
    }
    int prLo = 0, prHi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prLo, &prHi);   // prHi is the numerically smallest
    for (int b = 0; b < tpt_context::kMaxBands && e == cudaSuccess; ++b)
    {
        int pr = prHi + b; if (pr > prLo) pr = prLo;                              // earlier bands first
        e = cudaStreamCreateWithPriority(&ctx->bandStream[b], cudaStreamNonBlocking, pr);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->bandEvent[b], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { tpt_destroy(ctx); return (int)e; }
    *out = ctx;
    return 0;
}

void tpt_destroy(tpt_context* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->registeredPtr) cudaHostUnregister(ctx->registeredPtr);
    cudaFree(ctx->dImage); cudaFree(ctx->dScratch); cudaFree(ctx->dRayCounters); cudaFree(ctx->dLook); cudaFree(ctx->dLookRays);
    cudaFree(ctx->dAccum); cudaFree(ctx->dWork);
    for (int i = 0; i < 2; ++i)
    {
        cudaFree(ctx->dBlobs[i]);
        if (ctx->hBlob[i]) cudaFreeHost(ctx->hBlob[i]);
        if (ctx->uploadDone[i]) cudaEventDestroy(ctx->uploadDone[i]);
        if (ctx->blobLastUse[i]) cudaEventDestroy(ctx->blobLastUse[i]);
    }
    if (ctx->uploadStream) cudaStreamDestroy(ctx->uploadStream);
    if (ctx->lastDraw) cudaEventDestroy(ctx->lastDraw);
    if (ctx->tlKernelEnd) cudaEventDestroy(ctx->tlKernelEnd);
    for (int b = 0; b < 16; ++b) if (ctx->tlBand[b]) cudaEventDestroy(ctx->tlBand[b]);
    if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
    for (int b = 0; b < tpt_context::kMaxBands; ++b)
    {
        if (ctx->bandStream[b]) cudaStreamDestroy(ctx->bandStream[b]);
        if (ctx->bandEvent[b]) cudaEventDestroy(ctx->bandEvent[b]);
    }
    if (ctx->forkEvent) cudaEventDestroy(ctx->forkEvent);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    cudaFree(ctx->dBandDone);
    if (ctx->evStart) cudaEventDestroy(ctx->evStart);
    if (ctx->evStop) cudaEventDestroy(ctx->evStop);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* tpt_last_error(tpt_context* ctx) { return ctx ? ctx->lastError.c_str() : "null context"; }

int tpt_set_scene(tpt_context* ctx, const void* spheres20, const void* materials36, int count,
                  const void* camera88, const int* emissives, int emissiveCount)
{
    if (!ctx) return (int)cudaErrorInvalidValue;
    if (!spheres20 || !materials36 || !camera88 || count <= 0) return fail_msg(ctx, "tpt_set_scene: bad arguments");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    // the caller's emissive list indexes spheres/materials (Test.cpp:96-100): refuse what would read out of bounds
    if (emissives)
    {
        if (emissiveCount < 0 || emissiveCount > count) return fail_msg(ctx, "tpt_set_scene: emissiveCount out of range");
        for (int i = 0; i < emissiveCount; ++i)
            if (emissives[i] < 0 || emissives[i] >= count) return fail_msg(ctx, "tpt_set_scene: emissive id out of range");
    }
    std::vector<unsigned char> blob;
    SceneBlobLayout L;
    int nLights = 0;
    pack_scene_blob((const Sphere20*)spheres20, (const Material36*)materials36, count, emissives, emissiveCount, blob, L, nLights,
                    ctx->sceneFlags);
    // geometry + light table must fit in shared memory next to the kernels' static shared arrays
    const uint32_t kMaxStage = 200 * 1024;
    if (L.geomBytes > kMaxStage) return fail_msg(ctx, "tpt_set_scene: too many spheres for shared-memory staging");
    // A reference shell calls UpdateTest + upload every frame although the scene only changes under kFlagAnimate
    // (Test.cpp:304-308): identical bytes are not uploaded again.
    const bool same = ctx->haveScene && ctx->scene.count == count && ctx->scene.layout.flags == L.flags && blob == ctx->lastBlob;
    ctx->lastSceneUploadBytes = 0;
    if (!same || ctx->sceneUploadAlways)
    {
        const int slot = ctx->haveScene ? (ctx->curBlob ^ 1) : 0;
        // the slot's previous contents may still be read by draws issued before the last upload (device side: the
        // upload stream waits on the slot's last-use event) and its staging buffer by that upload's DMA (host side)
        CK(cudaEventSynchronize(ctx->uploadDone[slot]), "sync staging slot");
        if (blob.size() > ctx->blobCap[slot])
        {
            CK(cudaEventSynchronize(ctx->blobLastUse[slot]), "sync before scene realloc");
            cudaFree(ctx->dBlobs[slot]); ctx->dBlobs[slot] = nullptr; ctx->blobCap[slot] = 0;
            CK(cudaMalloc(&ctx->dBlobs[slot], blob.size()), "cudaMalloc scene");
            ctx->blobCap[slot] = blob.size();
        }
        if (blob.size() > ctx->hBlobCap[slot])
        {
            if (ctx->hBlob[slot]) cudaFreeHost(ctx->hBlob[slot]);
            ctx->hBlob[slot] = nullptr; ctx->hBlobCap[slot] = 0;
            CK(cudaMallocHost(&ctx->hBlob[slot], blob.size()), "cudaMallocHost scene staging");
            ctx->hBlobCap[slot] = blob.size();
        }
        memcpy(ctx->hBlob[slot], blob.data(), blob.size());
        CK(cudaStreamWaitEvent(ctx->uploadStream, ctx->blobLastUse[slot], 0), "upload waits for the slot's readers");
        CK(cudaMemcpyAsync(ctx->dBlobs[slot], ctx->hBlob[slot], blob.size(), cudaMemcpyHostToDevice, ctx->uploadStream), "scene upload");
        CK(cudaEventRecord(ctx->uploadDone[slot], ctx->uploadStream), "upload event");
        ctx->curBlob = slot;
        ctx->lastSceneUploadBytes = (long long)blob.size();
        ctx->lastBlob.swap(blob);
        if (!same) ++ctx->sceneGen;              // a forced re-upload of identical bytes is not a new scene (frame-lookahead key)
    }
    ctx->scene.blob = ctx->dBlobs[ctx->curBlob];
    ctx->scene.layout = L;
    ctx->scene.count = count;
    ctx->scene.nLights = nLights;
    // small scenes: stage everything (spheres + materials); large ones: geometry only, materials from L2
    ctx->scene.stagedBytes = L.totalBytes <= 64 * 1024 ? L.totalBytes : L.geomBytes;
    // Expanded-form sweep gate (FastHitterK, tpt_fast.cu). With u = 2^-24 the expanded form evaluates
    // c = |s-o|^2 - r^2 with |dc| <= ~4u (|s| + |o|)^2, the reference form (Maths.cpp:97-102) with |dc| <= ~4u |s-o|^2 + u r^2.
    // The gate |s|^2 <= 128 + 2 r^2 keeps (|s| + |o|)^2 within a small multiple of max(r^2, 128) for origins inside the
    // scene's extent (|o|^2 <= 128), i.e. the expanded form is no worse than ~8x the reference form's own bound for small
    // spheres and BETTER than it for huge ones (ground sphere r = 100: 2 s.o ~ 400 vs |co|^2 ~ 10^4). The bound is not
    // what the parity claim rests on: tests/test_gpu_fast.py compares kform 0 and 1 against the bit-exact mode at
    // >= 16k spp, overall and per first-hit material.
    ctx->scene.kformOk = true;
    for (int i = 0; i < count; ++i)
    {
        const Sphere20& s = ((const Sphere20*)spheres20)[i];
        const double c2 = (double)s.center[0] * s.center[0] + (double)s.center[1] * s.center[1] + (double)s.center[2] * s.center[2];
        if (c2 > 128.0 + 2.0 * (double)s.radius * s.radius) ctx->scene.kformOk = false;
    }
    memcpy(&ctx->cam, camera88, sizeof(Camera88));
    ctx->haveScene = true;
    return 0;
}

int tpt_set_camera(tpt_context* ctx, const void* camera88)
{
    if (!ctx || !camera88) return (int)cudaErrorInvalidValue;
    memcpy(&ctx->cam, camera88, sizeof(Camera88));
    return 0;
}

int tpt_set_spp(tpt_context* ctx, int spp)
{
    if (!ctx) return (int)cudaErrorInvalidValue;
    if (spp < 1 || spp > 4096) return fail_msg(ctx, "tpt_set_spp: spp out of range");
    ctx->spp = spp;
    return 0;
}

int tpt_set_option(tpt_context* ctx, const char* key, int value)
{
    if (!ctx || !key) return (int)cudaErrorInvalidValue;
    if (!strcmp(key, "fast_variant")) { if (value < -1 || value > 9) return fail_msg(ctx, "fast_variant: -1 (auto), 0..9"); ctx->fastVariant = value; return 0; }
    if (!strcmp(key, "exact_lanes")) { if (value != 0 && value != 1 && value != 2 && value != 8 && value != 9 && value != 32 && (value < 64 || value > 71)) return fail_msg(ctx, "exact_lanes: 0,1,2,8,9,32,64..71"); ctx->exactLanes = value; return 0; }
    if (!strcmp(key, "exact_lookahead")) { if (value < -1 || value > 256) return fail_msg(ctx, "exact_lookahead: -1 (adaptive), 0..256"); ctx->exactLookahead = value; ctx->lookValid = false; ctx->lookKeyValid = false; return 0; }
    if (!strcmp(key, "register_host")) { ctx->registerHost = value ? 1 : 0; return 0; }
    if (!strcmp(key, "fast_kform")) { if (value < 0 || value > 2) return fail_msg(ctx, "fast_kform: 0..2"); ctx->fastKForm = value; return 0; }
    if (!strcmp(key, "fast_alpha_zero")) { ctx->fastAlphaZero = value ? 1 : 0; return 0; }
    if (!strcmp(key, "scene_upload_always")) { ctx->sceneUploadAlways = value ? 1 : 0; return 0; }
    if (!strcmp(key, "mitsuba_compare")) { ctx->sceneFlags = value ? (ctx->sceneFlags | kSceneMitsuba) : (ctx->sceneFlags & ~(uint32_t)kSceneMitsuba); return 0; }
    if (!strcmp(key, "host_progress")) { ctx->hostProgress = value ? 1 : 0; return 0; }
    if (!strcmp(key, "host_zero_copy")) { ctx->hostZeroCopy = value ? 1 : 0; return 0; }
    if (!strcmp(key, "progress_bands")) { if (value < 1 || value > 16) return fail_msg(ctx, "progress_bands: 1..16"); ctx->progressBands = value; return 0; }
    if (!strcmp(key, "host_bands")) { if (value < 1 || value > tpt_context::kMaxBands) return fail_msg(ctx, "host_bands: 1..8"); ctx->hostBands = value; return 0; }
    if (!strcmp(key, "max_scratch_mb")) { if (value < 16) return fail_msg(ctx, "max_scratch_mb: >= 16"); ctx->maxScratchBytes = (size_t)value << 20; return 0; }
    return fail_msg(ctx, "tpt_set_option: unknown key");
}

// Copies the rows y_i = row0 + i*rowStep (or the packed band) of a float4 image between host and device.
static cudaError_t copy_rows(float* dst, const float* src, int width, int row0, int numRows, int rowStep, int packed,
                             cudaMemcpyKind kind, cudaStream_t stream)
{
    const size_t rowBytes = (size_t)width * 16;
    if (packed) return cudaMemcpyAsync(dst, src, rowBytes * numRows, kind, stream);
    const size_t off = (size_t)row0 * width * 4;
    if (rowStep == 1) return cudaMemcpyAsync(dst + off, src + off, rowBytes * numRows, kind, stream);
    return cudaMemcpy2DAsync(dst + off, rowBytes * rowStep, src + off, rowBytes * rowStep, rowBytes, numRows, kind, stream);
}

static int ensure(tpt_context* ctx, void** p, size_t* cap, size_t bytes, const char* what)
{
    if (bytes <= *cap) return 0;
    CK(cudaStreamSynchronize(ctx->stream), "sync before realloc");
    cudaFree(*p); *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, bytes), what);
    *cap = bytes;
    return 0;
}

int tpt_draw(tpt_context* ctx, int frameCount, int numFrames, int width, int height,
             int row0, int numRows, int rowStep, int packed,
             float* backbuffer, int bufferOnDevice, unsigned testFlags, int mode,
             long long* outRayCount, long long* outRaysPerFrame, void* cudaStreamArg)
{
    if (!ctx) return (int)cudaErrorInvalidValue;
    if (!ctx->haveScene) return fail_msg(ctx, "tpt_draw: no scene (call tpt_set_scene after UpdateTest)");
    if (!backbuffer || width <= 0 || height <= 0 || numFrames <= 0 || frameCount < 0 || numRows < 0 || rowStep <= 0 || row0 < 0)
        return fail_msg(ctx, "tpt_draw: bad arguments");
    if (mode != TPT_MODE_EXACT && mode != TPT_MODE_FAST && mode != TPT_MODE_REFGPU && mode != TPT_MODE_REFGPU_FAST)
        return fail_msg(ctx, "tpt_draw: unknown mode");
    const bool refgpu = mode == TPT_MODE_REFGPU || mode == TPT_MODE_REFGPU_FAST;
    if (numRows == 0)
    {
        // an empty shard (more ranks than rows, TraceRowJob(start, start)): nothing to trace, nothing to copy
        if (outRayCount) *outRayCount = 0;
        if (outRaysPerFrame) for (int i = 0; i < numFrames; ++i) outRaysPerFrame[i] = 0;
        ctx->lastLaunches = 0;
        return 0;
    }
    if (row0 + (long long)(numRows - 1) * rowStep >= height) return fail_msg(ctx, "tpt_draw: rows outside the image");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t stream = cudaStreamArg ? (cudaStream_t)cudaStreamArg : ctx->stream;
    // One context = one scene blob, one set of counters: draws are ordered in issue order even across streams, and the
    // scene this draw reads must have landed (tpt_set_scene uploads asynchronously on its own stream).
    if (ctx->haveLastDraw && ctx->lastDrawStream != stream) CK(cudaStreamWaitEvent(stream, ctx->lastDraw, 0), "order after previous draw");
    CK(cudaStreamWaitEvent(stream, ctx->uploadDone[ctx->curBlob], 0), "wait for scene upload");
    SceneDev scene = ctx->scene;
    scene.kformOk = scene.kformOk && ctx->fastKForm != 0;
    scene.kformMode = ctx->fastKForm;

    const size_t bufRows = packed ? (size_t)numRows : (size_t)height;
    const size_t bufBytes = bufRows * width * 4 * sizeof(float);
    float* dImage = backbuffer;
    bool needPrev = true;
    bool zeroCopy = false;
    if (!bufferOnDevice && ctx->registerHost && (ctx->registeredPtr != backbuffer || ctx->registeredBytes != bufBytes))
    {
        if (ctx->registeredPtr) { cudaHostUnregister(ctx->registeredPtr); ctx->registeredPtr = nullptr; }
        // page-lock (and map) the caller's buffer once: copies run at full PCIe rate and the kernels can store into it
        // directly; failure is not fatal
        if (cudaHostRegister(backbuffer, bufBytes, cudaHostRegisterMapped) == cudaSuccess)
        {
            ctx->registeredPtr = backbuffer; ctx->registeredBytes = bufBytes;
        }
        else (void)cudaGetLastError();
    }
    // auto: the warp slab queue. Only where the sweep dwarfs everything else (>= 1024 spheres) AND the queue kernel cannot
    // run a packed-pair sweep (pair array does not fit, or "fast_kform" < 2) is the block wavefront that sorts its paths
    // by material before Scatter() faster (4096 spheres, Mray/s: queue packed 443, wavefront 333, queue scalar 312)
    const bool waveFits = (long long)((numRows * (long long)width + 127) / 128) * 128 * ctx->spp * (numFrames < 256 ? numFrames : 256) <= 0x7fffffffLL;
    SceneDev autoScene = ctx->scene;
    autoScene.kformMode = ctx->fastKForm;
    int fastVariant = ctx->fastVariant < 0 ? (ctx->scene.count >= 1024 && waveFits && fast_queue_kform(autoScene) < 2 ? 7 : 3) : ctx->fastVariant;
    if (!bufferOnDevice && mode == TPT_MODE_FAST && ctx->hostZeroCopy && (ctx->fastVariant < 0 || fast_variant_writes_final_pixels(ctx->fastVariant)))
    {
        // Host-buffer draw whose `prev` has zero weight, into page-locked memory the GPU can address: the trace kernel's
        // coalesced 128-bit pixel stores go straight into the caller's buffer over PCIe while the other pixels are still
        // being traced — no staging image, no device-to-host copy after the kernel.
        float wPrev = 1.0f;
        for (int f = 0; f < numFrames; ++f) wPrev *= lerp_fac(frameCount + f, testFlags);
        cudaPointerAttributes attr;
        if (wPrev == 0.0f && numFrames <= 256 && cudaPointerGetAttributes(&attr, backbuffer) == cudaSuccess &&
            attr.type == cudaMemoryTypeHost && attr.devicePointer)
        {
            dImage = (float*)attr.devicePointer;
            zeroCopy = true;
            needPrev = false;
            fastVariant = ctx->fastVariant < 0 ? 8 : ctx->fastVariant;
        }
        else (void)cudaGetLastError();
    }
    if (!bufferOnDevice && !zeroCopy)
    {
        int r = ensure(ctx, (void**)&ctx->dImage, &ctx->imageCap, bufBytes, "cudaMalloc image");
        if (r) return r;
        dImage = ctx->dImage;
        // `prev` is an input of the blend (Test.cpp:293). Exact mode always uploads it (NaN/Inf * 0 and the
        // untouched alpha are part of bit parity). Fast mode uploads it only when prev has a non-zero weight;
        // otherwise the kernel writes alpha = 0, which is what every reference shell's zero-initialised buffer
        // holds (TestWin.cpp:73-74, Renderer.mm:148-149, Emscripten/main.cpp:50-51).
        needPrev = mode == TPT_MODE_EXACT;
        if (!needPrev)
        {
            // (REFGPU writes alpha = 1 and lerp(col, prev, 0) does not read prev: same rule as the fast mode)
            float wPrev = 1.0f;
            for (int f = 0; f < numFrames; ++f) wPrev *= lerp_fac(frameCount + f, testFlags);
            needPrev = wPrev != 0.0f;
        }
        if (needPrev)
        {
            cudaError_t e = copy_rows(dImage, backbuffer, width, row0, numRows, rowStep, packed, cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) return fail(ctx, e, "H2D backbuffer");
        }
    }

    // frames per launch: exact mode needs numFrames*numRows*width float4 of scratch when numFrames > 1
    int framesPerLaunch = numFrames;
    if (mode == TPT_MODE_EXACT && numFrames > 1)
    {
        const size_t perFrame = (size_t)numRows * width * 16;
        size_t fit = ctx->maxScratchBytes / perFrame;
        if (fit < 1) fit = 1;
        if ((size_t)framesPerLaunch > fit) framesPerLaunch = (int)fit;
    }
    if (mode == TPT_MODE_FAST && framesPerLaunch > 256) framesPerLaunch = 256;


    if (numFrames > ctx->rayCounterCap)
    {
        CK(cudaStreamSynchronize(stream), "sync before counter realloc");
        cudaFree(ctx->dRayCounters); ctx->dRayCounters = nullptr; ctx->rayCounterCap = 0;
        CK(cudaMalloc(&ctx->dRayCounters, (size_t)numFrames * sizeof(unsigned long long)), "cudaMalloc counters");
        ctx->rayCounterCap = numFrames;
    }
    CK(cudaMemsetAsync(ctx->dRayCounters, 0, (size_t)numFrames * sizeof(unsigned long long), stream), "zero counters");

    DrawParams p;
    memset(&p, 0, sizeof(p));
    p.cam = ctx->cam;
    p.width = width; p.height = height;
    p.row0 = row0; p.numRows = numRows; p.rowStep = rowStep; p.packed = packed ? 1 : 0;
    p.spp = ctx->spp;
    p.flags = testFlags;
    p.invWidth = 1.0f / (float)width;     // Test.cpp:270-271
    p.invHeight = 1.0f / (float)height;
    p.image = dImage;
    p.workCounter = ctx->dWork + (size_t)(ctx->workSlot++ % tpt_context::kWorkSlots) * 32;
    // alpha is never written by the reference (Maths.h:38). When `prev` has zero weight the fast kernels keep the alpha
    // that is in the buffer, unless the staging image was not uploaded (host buffer) or the caller waived it.
    p.zeroAlpha = (ctx->fastAlphaZero || (!bufferOnDevice && !needPrev)) ? 1 : 0;     // (zero-copy: never reads host memory)

    ctx->lastLaunches = 0;
    CK(cudaEventRecord(ctx->evStart, stream), "event record");
    // Host-buffer fast draws: split the rows into bands, one stream per band (earlier band = higher priority). Each
    // stream runs prepare + trace for its band and then copies the band to the caller's buffer, so the D2H of band b
    // overlaps the tracing of band b+1 and the persistent CTAs of band b+1 fill the SMs as band b's tail drains.
    bool pipelined = mode == TPT_MODE_FAST && !bufferOnDevice && (fastVariant == 9 || (fastVariant >= 3 && fastVariant <= 7)) && ctx->hostBands > 1 &&
                     (rowStep == 1 || packed) && framesPerLaunch == numFrames && numRows >= 16 * ctx->hostBands;
    // the per-band completion counters are 32-bit (cuStreamWaitValue32): a band never holds more paths than the image
    const bool progress = pipelined && ctx->waitValue32 && ctx->hostProgress && (fastVariant == 3 || fastVariant == 4 || fastVariant == 9) &&
                          (long long)numRows * width * ctx->spp * numFrames <= 0xFFFFFFFFLL;
    if (progress)
    {
        // ONE kernel for the whole image; the copy stream waits on the kernel's per-band completion counters
        const int NB = ctx->progressBands;
        unsigned int expected[16];
        CK(cudaMemsetAsync(ctx->dBandDone, 0, 64, stream), "zero band counters");
        CK(cudaEventRecord(ctx->forkEvent, stream), "fork event");
        CK(cudaStreamWaitEvent(ctx->copyStream, ctx->forkEvent, 0), "copy stream wait");
        p.frame0 = frameCount; p.numFrames = numFrames; p.rayCounter = ctx->dRayCounters;
        cudaError_t e = launch_fast(p, scene, fastVariant, ctx->numSMs, stream, ctx->dBandDone, NB, expected);
        if (e != cudaSuccess) return fail(ctx, e, "kernel launch");
        ctx->lastLaunches += fast_kernel_launches(p, fastVariant);
        CK(cudaEventRecord(ctx->tlKernelEnd, stream), "timeline event");
        ctx->tlBands = 0;
        const long long regionPix = (long long)numRows * width, slab = fast_slab_pixels();
        const long long mtiles = (regionPix + slab - 1) / slab, mpb = (mtiles + NB - 1) / NB;
        const size_t firstPix = packed ? 0 : (size_t)row0 * width;
        for (int b = NB - 1; b >= 0; --b)     // the kernel walks the image from its last macro-tile down: last band first
        {
            const long long p0 = (long long)b * mpb * slab, p1 = (long long)(b + 1) * mpb * slab < regionPix ? (long long)(b + 1) * mpb * slab : regionPix;
            if (p1 <= p0) continue;
            if (ctx->waitValue32((CUstream)ctx->copyStream, (CUdeviceptr)(ctx->dBandDone + b), expected[b], CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                return fail_msg(ctx, "cuStreamWaitValue32 failed");
            const size_t off = (firstPix + (size_t)p0) * 4;
            CK(cudaMemcpyAsync(backbuffer + off, dImage + off, (size_t)(p1 - p0) * 16, cudaMemcpyDeviceToHost, ctx->copyStream), "D2H band");
            CK(cudaEventRecord(ctx->tlBand[ctx->tlBands++], ctx->copyStream), "timeline event");
        }
        CK(cudaEventRecord(ctx->bandEvent[0], ctx->copyStream), "copy event");
        CK(cudaStreamWaitEvent(stream, ctx->bandEvent[0], 0), "join copies");
    }
    else if (pipelined)
    {
        const int NB = ctx->hostBands;
        CK(cudaEventRecord(ctx->forkEvent, stream), "fork event");
        p.frame0 = frameCount; p.numFrames = numFrames; p.rayCounter = ctx->dRayCounters;
        for (int b = 0; b < NB; ++b)
        {
            const int rb0 = (int)((long long)numRows * b / NB), rb1 = (int)((long long)numRows * (b + 1) / NB);
            cudaStream_t bs = ctx->bandStream[b];
            CK(cudaStreamWaitEvent(bs, ctx->forkEvent, 0), "band wait");
            DrawParams pb = p;
            pb.row0 = row0 + rb0 * rowStep;
            pb.numRows = rb1 - rb0;
            pb.workCounter = p.workCounter + 4 * b;
            const size_t firstRow = packed ? (size_t)rb0 : (size_t)(row0 + rb0);   // rowStep == 1 when not packed
            if (packed) pb.image = dImage + firstRow * width * 4;
            cudaError_t e = launch_fast(pb, scene, fastVariant, ctx->numSMs, bs);
            if (e != cudaSuccess) return fail(ctx, e, "kernel launch");
            ctx->lastLaunches += fast_kernel_launches(pb, fastVariant);
            const size_t off = firstRow * width * 4, bytes = (size_t)(rb1 - rb0) * width * 16;
            CK(cudaMemcpyAsync(backbuffer + off, dImage + off, bytes, cudaMemcpyDeviceToHost, bs), "D2H band");
            CK(cudaEventRecord(ctx->bandEvent[b], bs), "band event");
            CK(cudaStreamWaitEvent(stream, ctx->bandEvent[b], 0), "join band");
        }
    }
    for (int f = 0; f < numFrames && !pipelined; f += framesPerLaunch)
    {
        const int nf = numFrames - f < framesPerLaunch ? numFrames - f : framesPerLaunch;
        p.frame0 = frameCount + f;
        p.numFrames = nf;
        cudaError_t e;
        if (mode == TPT_MODE_EXACT && numFrames == 1 && (ctx->exactLookahead > 1 || ctx->exactLookahead == -1) && !(testFlags & TPT_FLAG_ANIMATE))
        {
            // frame lookahead (see tpt_context::exactLookahead)
            tpt_context::LookKey k{};
            k.width = width; k.height = height; k.row0 = row0; k.numRows = numRows; k.rowStep = rowStep; k.spp = ctx->spp;
            k.sceneGen = ctx->sceneGen; k.cam = ctx->cam;
            const tpt_context::LookKey& c = ctx->look;
            const size_t perFrame = (size_t)numRows * width * 4;           // floats
            const bool sameSetup = c.width == k.width && c.height == k.height && c.row0 == k.row0 && c.numRows == k.numRows &&
                                   c.rowStep == k.rowStep && c.spp == k.spp && c.sceneGen == k.sceneGen && !memcmp(&c.cam, &k.cam, sizeof(Camera88));
            const bool hit = ctx->lookValid && frameCount >= c.frame0 && frameCount < c.frame0 + c.n && sameSetup;
            if (!hit)
            {
                size_t L = ctx->exactLookahead > 1 ? (size_t)ctx->exactLookahead : 1;
                if (ctx->exactLookahead == -1 && ctx->lookKeyValid && sameSetup && frameCount == c.frame0 + c.n &&
                    ctx->lookLastServed == c.frame0 + c.n - 1)
                    L = c.n >= 8 ? 16 : (size_t)c.n * 2;           // the caller consumed the whole window and continues: double it
                const size_t fit = ctx->maxScratchBytes / (perFrame * 4);
                if (L > fit) L = fit < 1 ? 1 : fit;
                int r = ensure(ctx, (void**)&ctx->dLook, &ctx->lookCap, L * perFrame * 4, "cudaMalloc lookahead cache");
                if (r) return r;
                r = ensure(ctx, (void**)&ctx->dLookRays, &ctx->lookRaysCap, L * sizeof(unsigned long long), "cudaMalloc lookahead counters");
                if (r) return r;
                CK(cudaMemsetAsync(ctx->dLookRays, 0, L * sizeof(unsigned long long), stream), "zero lookahead counters");
                DrawParams pl = p;
                pl.frame0 = frameCount; pl.numFrames = (int)L; pl.scratch = ctx->dLook; pl.rayCounter = ctx->dLookRays;
                e = launch_exact(pl, scene, ctx->exactLanes, stream, /*resolve*/ L == 1);
                if (e != cudaSuccess) return fail(ctx, e, "kernel launch");
                ctx->lastLaunches += 1;
                k.frame0 = frameCount; k.n = (int)L;
                ctx->look = k; ctx->lookValid = L > 1; ctx->lookKeyValid = true; ctx->lookLastServed = frameCount;
                if (L == 1) { CK(cudaMemcpyAsync(ctx->dRayCounters, ctx->dLookRays, 8, cudaMemcpyDeviceToDevice, stream), "counter copy"); continue; }
            }
            const int i = frameCount - ctx->look.frame0;
            ctx->lookLastServed = frameCount;
            DrawParams pr = p;
            pr.frame0 = frameCount; pr.numFrames = 1; pr.scratch = ctx->dLook + (size_t)i * perFrame;
            e = launch_resolve_exact(pr, stream);
            ctx->lastLaunches += 1;
            CK(cudaMemcpyAsync(ctx->dRayCounters, ctx->dLookRays + i, 8, cudaMemcpyDeviceToDevice, stream), "counter copy");
        }
        else if (mode == TPT_MODE_EXACT)
        {
            p.rayCounter = ctx->dRayCounters + f;
            p.scratch = nullptr;
            if (nf > 1)
            {
                int r = ensure(ctx, (void**)&ctx->dScratch, &ctx->scratchCap, (size_t)nf * numRows * width * 16, "cudaMalloc scratch");
                if (r) return r;
                p.scratch = ctx->dScratch;
            }
            e = launch_exact(p, scene, ctx->exactLanes, stream);
            ctx->lastLaunches += nf > 1 ? 2 : 1;
        }
        else if (refgpu)
        {
            p.rayCounter = ctx->dRayCounters;
            e = mode == TPT_MODE_REFGPU ? launch_refgpu_exact(p, scene, ctx->numSMs, stream) : launch_refgpu_fast(p, scene, ctx->numSMs, stream);
            ctx->lastLaunches += 1;
        }
        else
        {
            p.rayCounter = ctx->dRayCounters;
            e = launch_fast(p, scene, fastVariant, ctx->numSMs, stream);
            ctx->lastLaunches += fast_kernel_launches(p, fastVariant);
        }
        if (e != cudaSuccess) return fail(ctx, e, "kernel launch");
    }
    CK(cudaEventRecord(ctx->evStop, stream), "event record");
    ctx->haveTiming = true;
    const bool wantCounts = outRayCount || outRaysPerFrame;
    if (wantCounts && numFrames > ctx->hPinnedCap)
    {
        if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
        ctx->hPinned = nullptr; ctx->hPinnedCap = 0;
        CK(cudaMallocHost(&ctx->hPinned, (size_t)numFrames * sizeof(unsigned long long)), "cudaMallocHost");
        ctx->hPinnedCap = numFrames;
    }
    k_accumulate_rays<<<1, 1, 0, stream>>>(ctx->dRayCounters, numFrames, ctx->dAccum, wantCounts ? ctx->hPinned : nullptr);
    CK(cudaGetLastError(), "accumulate launch");
    ctx->lastLaunches += 1;

    if (!bufferOnDevice && !pipelined && !zeroCopy)
    {
        // only the rows this draw rendered go back: the caller's other rows are not ours to touch (TraceRowJob writes
        // rows [start,end) only, Test.cpp:278-297)
        cudaError_t e = copy_rows(backbuffer, dImage, width, row0, numRows, rowStep, packed, cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return fail(ctx, e, "D2H backbuffer");
    }
    CK(cudaEventRecord(ctx->lastDraw, stream), "draw end event");
    CK(cudaEventRecord(ctx->blobLastUse[ctx->curBlob], stream), "scene last-use event");
    ctx->lastDrawStream = stream; ctx->haveLastDraw = true;

    if (outRayCount || outRaysPerFrame)
    {
        CK(cudaStreamSynchronize(stream), "stream sync");     // the counts were stored into hPinned by k_accumulate_rays
        long long total = 0;
        for (int i = 0; i < numFrames; ++i)
        {
            total += (long long)ctx->hPinned[i];
            if (outRaysPerFrame) outRaysPerFrame[i] = (long long)ctx->hPinned[i];
        }
        if (outRayCount) *outRayCount = total;
    }
    else if (!bufferOnDevice)
        CK(cudaStreamSynchronize(stream), "stream sync"); // host-buffer draws are synchronous like DrawTest
    return 0;
}

int tpt_read_ray_count(tpt_context* ctx, void* cudaStreamArg, long long* outRays)
{
    if (!ctx || !outRays) return (int)cudaErrorInvalidValue;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t stream = cudaStreamArg ? (cudaStream_t)cudaStreamArg : ctx->stream;
    unsigned long long h[2] = {0, 0};
    CK(cudaMemcpyAsync(h, ctx->dAccum, sizeof(h), cudaMemcpyDeviceToHost, stream), "D2H accum");
    CK(cudaMemsetAsync(ctx->dAccum, 0, sizeof(unsigned long long), stream), "reset accum");
    CK(cudaStreamSynchronize(stream), "stream sync");
    *outRays = (long long)h[0];
    return 0;
}

int tpt_last_kernel_ms(tpt_context* ctx, float* outMs)
{
    if (!ctx || !outMs) return (int)cudaErrorInvalidValue;
    if (!ctx->haveTiming) return fail_msg(ctx, "tpt_last_kernel_ms: no draw yet");
    CK(cudaEventSynchronize(ctx->evStop), "event sync");
    CK(cudaEventElapsedTime(outMs, ctx->evStart, ctx->evStop), "event elapsed");
    return 0;
}

int tpt_last_launch_count(tpt_context* ctx) { return ctx ? ctx->lastLaunches : 0; }
long long tpt_last_scene_upload_bytes(tpt_context* ctx) { return ctx ? ctx->lastSceneUploadBytes : 0; }

int tpt_tonemap_rgba8(tpt_context* ctx, const float* image, int imageOnDevice, int width, int height,
                      unsigned char* dst, int dstOnDevice, int transfer, int bgr, int flipY, void* cudaStreamArg)
{
    if (!ctx || !image || !dst || width <= 0 || height <= 0) return (int)cudaErrorInvalidValue;
    if (transfer < 0 || transfer > 2) return fail_msg(ctx, "tpt_tonemap_rgba8: transfer 0..2");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaStream_t stream = cudaStreamArg ? (cudaStream_t)cudaStreamArg : ctx->stream;
    const size_t inBytes = (size_t)width * height * 16, outBytes = (size_t)width * height * 4;
    const float* dIn = image;
    if (!imageOnDevice)
    {
        int r = ensure(ctx, (void**)&ctx->dImage, &ctx->imageCap, inBytes, "cudaMalloc image");
        if (r) return r;
        CK(cudaMemcpyAsync(ctx->dImage, image, inBytes, cudaMemcpyHostToDevice, stream), "H2D image");
        dIn = ctx->dImage;
    }
    unsigned char* dOut = dst;
    if (!dstOnDevice)
    {
        int r = ensure(ctx, (void**)&ctx->dScratch, &ctx->scratchCap, outBytes, "cudaMalloc tonemap out");
        if (r) return r;
        dOut = (unsigned char*)ctx->dScratch;
    }
    const int n = width * height;
    k_tonemap_rgba8<<<(n + 255) / 256, 256, 0, stream>>>((const float4*)dIn, width, height, (uchar4*)dOut, transfer, bgr ? 1 : 0, flipY ? 1 : 0);
    CK(cudaGetLastError(), "tonemap launch");
    if (!dstOnDevice)
    {
        CK(cudaMemcpyAsync(dst, dOut, outBytes, cudaMemcpyDeviceToHost, stream), "D2H tonemap");
        CK(cudaStreamSynchronize(stream), "stream sync");
    }
    return 0;
}

int tpt_tonemap_srgb8(tpt_context* ctx, const float* image, int imageOnDevice, int width, int height,
                      unsigned char* dst, int dstOnDevice, void* cudaStreamArg)
{
    return tpt_tonemap_rgba8(ctx, image, imageOnDevice, width, height, dst, dstOnDevice, 0, 0, 1, cudaStreamArg);
}

int tpt_mem_alloc(tpt_context* ctx, unsigned long long bytes, void** outDevPtr)
{
    if (!ctx || !outDevPtr || !bytes) return (int)cudaErrorInvalidValue;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaMalloc(outDevPtr, (size_t)bytes), "cudaMalloc");
    CK(cudaMemset(*outDevPtr, 0, (size_t)bytes), "cudaMemset");
    return 0;
}
int tpt_mem_free(tpt_context* ctx, void* devPtr)
{
    if (!ctx) return (int)cudaErrorInvalidValue;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaFree(devPtr), "cudaFree");
    return 0;
}
int tpt_mem_copy(tpt_context* ctx, void* dst, const void* src, unsigned long long bytes, int kind)
{
    if (!ctx || !dst || !src || kind < 1 || kind > 3) return (int)cudaErrorInvalidValue;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaStreamSynchronize(ctx->stream), "stream sync");
    CK(cudaMemcpy(dst, src, (size_t)bytes, kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice), "cudaMemcpy");
    return 0;
}
int tpt_ipc_export(tpt_context* ctx, void* devPtr, void* outHandle64)
{
    if (!ctx || !devPtr || !outHandle64) return (int)cudaErrorInvalidValue;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, devPtr), "cudaIpcGetMemHandle");
    memcpy(outHandle64, &h, 64);
    return 0;
}
int tpt_ipc_open(tpt_context* ctx, const void* handle64, void** outDevPtr)
{
    if (!ctx || !handle64 || !outDevPtr) return (int)cudaErrorInvalidValue;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(outDevPtr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return 0;
}
int tpt_ipc_close(tpt_context* ctx, void* devPtr)
{
    if (!ctx || !devPtr) return (int)cudaErrorInvalidValue;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    CK(cudaIpcCloseMemHandle(devPtr), "cudaIpcCloseMemHandle");
    return 0;
}

int tpt_debug_timeline(tpt_context* ctx, float* outMs, int capacity)
{
    if (!ctx || !outMs || capacity < 1) return (int)cudaErrorInvalidValue;
    CK(cudaEventSynchronize(ctx->evStop), "event sync");
    int n = 0;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->evStart, ctx->tlKernelEnd), "elapsed");
    outMs[n++] = ms;
    for (int b = 0; b < ctx->tlBands && n < capacity; ++b)
    {
        CK(cudaEventElapsedTime(&ms, ctx->evStart, ctx->tlBand[b]), "elapsed");
        outMs[n++] = ms;
    }
    if (n < capacity) { CK(cudaEventElapsedTime(&ms, ctx->evStart, ctx->evStop), "elapsed"); outMs[n++] = ms; }
    return -n;   // negative count = number of entries written
}

int tpt_debug_libm(tpt_context* ctx, int fn, const float* in, float* out, long long n)
{
    if (!ctx || !in || !out || n <= 0 || fn < 0 || fn > 3) return (int)cudaErrorInvalidValue;
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    float *dIn = nullptr, *dOut = nullptr;
    CK(cudaMalloc(&dIn, (size_t)n * 4), "cudaMalloc");
    cudaError_t e = cudaMalloc(&dOut, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dIn, in, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_debug_libm(fn, dIn, dOut, n, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dOut, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dIn); cudaFree(dOut);
    if (e != cudaSuccess) return fail(ctx, e, "tpt_debug_libm");
    return 0;
}

int tpt_debug_hit(tpt_context* ctx, int kform, const float* rays, int* outId, float* outT, long long n)
{
    if (!ctx || !rays || !outId || !outT || n <= 0 || kform < 0 || kform > 3) return (int)cudaErrorInvalidValue;
    if (!ctx->scene.blob) return fail_msg(ctx, "tpt_debug_hit: no scene");
    CK(cudaSetDevice(ctx->device), "cudaSetDevice");
    float *dRays = nullptr, *dT = nullptr;
    int* dId = nullptr;
    cudaError_t e = cudaMalloc(&dRays, (size_t)n * 24);
    if (e == cudaSuccess) e = cudaMalloc(&dT, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dId, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->uploadDone[ctx->curBlob], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dRays, rays, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_debug_hit(ctx->scene, kform, dRays, dId, dT, n, ctx->numSMs, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(outId, dId, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(outT, dT, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dRays); cudaFree(dT); cudaFree(dId);
    if (e != cudaSuccess) return fail(ctx, e, "tpt_debug_hit");
    return 0;
}

} // extern "C"
