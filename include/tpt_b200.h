/* tpt_b200 — C-ABI of the B200-native ToyPathTracer hot path.
 *
 * This library replaces what the reference's DrawTest() does on the CPU (Cpp/Source/Test.cpp:344-367 ->
 * TraceRowJob :266-300 -> Trace :195-234 -> HitWorld/HitSpheres Maths.cpp:50-203 -> Scatter :83-193) with
 * hand-written sm_100a CUDA kernels. It sits exactly where the reference's own GPU back-ends sit: the shell
 * calls UpdateTest(), pulls the scene with GetObjectCount()/GetSceneDesc() (raw 20 B Sphere / 36 B Material /
 * 88 B Camera structs + emissive id list, Cpp/Windows/TestWin.cpp:258-283) and dispatches a kernel instead of
 * DrawTest(). Plain pointers and sizes only; every function returns 0 on success or a CUDA error code
 * (tpt_last_error() gives the text). There is no CPU fallback: without a CUDA device tpt_create() fails.
 *
 * A drop-in C++ translation unit exporting the six functions of Cpp/Source/Test.h on top of this ABI is
 * toypathtracer_b200/csrc/test_shim.cpp.
 */
#ifndef TPT_B200_H
#define TPT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tpt_context tpt_context;

/* Rendering modes.
 * TPT_MODE_EXACT  bit-identical to the reference C++ path: per-(frame,row) XorShift32 streams (Test.cpp:280),
 *                 no FMA contraction, glibc-faithful sinf/cosf/powf, back-to-front colour fold. Same pixels,
 *                 same ray counts.
 * TPT_MODE_FAST   same estimator, one RNG stream per (pixel, sample, frame); statistically equivalent image and
 *                 rays/sample. Throughput mode. */
#define TPT_MODE_EXACT 0
#define TPT_MODE_FAST 1
/* TPT_MODE_REFGPU       the estimator of the reference's own GPU back-ends (Cpp/Windows/ComputeShader.hlsl:353-395 and the
 *                       Metal port): per-PIXEL seed (x*1973 + y*9277 + frames*26699)|1, analytic disk/sphere samplers,
 *                       saturating schlick, <= 10 path segments, lerp(col, prev, f) blend, alpha written as 1 — for like-
 *                       for-like comparison with the reference's published D3D11/Metal numbers (readme.md:64-77). Strict
 *                       IEEE arithmetic: bit-equal to the CPU restatement oracle/refgpu_restate.cpp.
 * TPT_MODE_REFGPU_FAST  the same with GPU-native arithmetic (FMA, MUFU), as a real shader compiler would emit. */
#define TPT_MODE_REFGPU 2
#define TPT_MODE_REFGPU_FAST 3

/* testFlags of Cpp/Source/Test.h:4-8 */
#define TPT_FLAG_ANIMATE 1u
#define TPT_FLAG_PROGRESSIVE 2u

/* Replaces InitializeTest() (Test.cpp:240-246, which only creates the CPU task scheduler): binds a context to
 * CUDA device `device`. */
int tpt_create(int device, tpt_context** out);
/* Replaces ShutdownTest() (Test.cpp:248-253). */
void tpt_destroy(tpt_context* ctx);
int tpt_device_count(void);
const char* tpt_last_error(tpt_context* ctx);

/* Takes exactly what GetSceneDesc() exports (Test.cpp:377-384): `count` 20 B spheres, `count` 36 B materials,
 * one 88 B camera, the emissive sphere ids (Test.cpp:321-338). emissives == NULL derives the list from the
 * materials the way UpdateTest does; ids outside [0,count) or a negative count are refused. invRadius is recomputed
 * (Maths.h:359). Call after every UpdateTest(). The call never blocks on the GPU: the packed scene goes through a
 * pinned staging buffer into the one of two device copies that no draw in flight reads (the per-frame
 * UpdateSubresource of Cpp/Windows/TestWin.cpp:261-283 as an async double-buffered upload); draws issued afterwards
 * wait for it on the device. Bytes identical to the scene already resident are not uploaded again. */
int tpt_set_scene(tpt_context* ctx, const void* spheres20, const void* materials36, int count,
                  const void* camera88, const int* emissives, int emissiveCount);
/* Camera only (UpdateTest rebuilds it every frame from the aspect ratio, Test.cpp:341). */
int tpt_set_camera(tpt_context* ctx, const void* camera88);

/* DO_SAMPLES_PER_PIXEL (Config.h:22), default 4. */
int tpt_set_spp(tpt_context* ctx, int spp);
/* Implementation knobs (benchmarks/tests): "fast_variant" (-1 auto (default): 3 for device buffers (7 from 1024 spheres when the packed pair array does not fit in shared memory), 8 for host-buffer
 * draws that can store straight into page-locked memory; 0 megakernel, 1/2 persistent tiles, 3/4 persistent slab queue with 128-bit L2
 * reductions (3 picks its 8-CTA/SM 64-register instance by itself for long draws; 9 forces that instance), 5 CTA-owned tiles,
 * 6/7 block wavefront with material sort, 8 warp-owned pixel groups with coalesced 128-bit write-out), "host_zero_copy" (default 1: with variant 8 a host-buffer draw whose `prev` has zero weight writes its finished
 * pixels directly into the caller's page-locked buffer over PCIe — no staging image, no device-to-host copy), "exact_lanes"
 * (0 auto; 64..67 = split kernel: one PATH warp per (frame,row) chain walks the RNG stream, 1..4 SHADE warps do the light
 * sampling / fold / blend off the critical path (auto for <= 2400 chains); 32 or 8 lanes per chain as nested loops; 1 = one
 * thread per chain as a flat one-sweep-per-step state machine (auto for >= 100 000 chains); 2 = one thread per chain
 * nested, 9 = 8 lanes flat: measured slower, kept for comparison; 68/69 timing probes; 70 = the split kernel as 2-CTA
 * clusters with the roles on different SMs and the rings over DSMEM, 71 = shade warps calling out-of-line libm: both measured
 * slower, kept for comparison), "register_host" (1: page-lock the caller's host backbuffer with cudaHostRegister the first
 * time it is seen so both copies run at full PCIe rate; only safe when the buffer outlives the context, as a
 * reference shell's does; default 0 — buffers that are already pinned are detected by CUDA on their own), "host_bands" (1..8, default 3:
 * host-buffer fast draws are split into row bands on separate streams so the D2H of one band overlaps the tracing of
 * the next), "host_progress" (default 1: with fast variant 3/4 a single kernel publishes per-band completion counters and the
 * copy stream waits on them with cuStreamWaitValue32, "progress_bands" bands, default 4; 0 falls back to host_bands),
 * "fast_kform" (fast kernels' sphere sweep: 0 reference form, 1 expanded form, 2 (default) expanded form evaluated two
 * spheres per instruction with Blackwell's packed fma.rn.f32x2; 1 and 2 only when the scene passes the gate in
 * tpt_set_scene — a scene that fails it gets, with 2, the packed sweep made conservative by a folded-in error bound
 * plus reference-form hit decisions, and the reference form otherwise; per context), "fast_alpha_zero" (default 0; 1: fast-mode draws whose `prev` has zero weight write
 * alpha = 0 instead of keeping the buffer's alpha — saves the read over NVLink when the buffer is a peer GPU's),
 * "exact_lookahead" (default 0; L > 1: an exact-mode draw of ONE frame that misses the cache traces frames [f, f+L) in a
 * single launch and keeps their per-frame colours; the calls for the following frames only blend their cached frame into
 * the caller's buffer. Bit-identical results and per-frame ray counts; trades L frames of latency on a miss for batch
 * throughput (one 720p frame alone cannot fill the GPU: `height` serial RNG chains). Never used with kFlagAnimate; any
 * scene / camera / size / row-range / spp change invalidates the cache. -1 = adaptive: the first call traces one frame,
 * and every time the caller has walked to the end of the cached window and asks for the frame right after it, the next
 * window doubles (1, 2, 4, 8, 16 frames) — no first-call latency; the Test.h shim sets this, TPT_EXACT_LOOKAHEAD overrides),
 * "scene_upload_always" (default 0; 1: tpt_set_scene uploads the scene blob even when its bytes equal the resident scene's —
 * bench.py's end-to-end leg, so that every step carries its host -> device copy),
 * "mitsuba_compare" (default 0; DO_MITSUBA_COMPARE of Config.h:25 as a runtime switch, applied by the NEXT
 * tpt_set_scene: constant sky (0.15, 0.21, 0.3) (Test.cpp:226-227) and zero Metal roughness (Test.cpp:143-145); the
 * switch's third effect, zero aperture (Test.cpp:312-313), is camera data: the Test.h shim's UpdateTest applies it). */
int tpt_set_option(tpt_context* ctx, const char* key, int value);

/* Replaces DrawTest() (Test.cpp:344-367) for frames [frameCount, frameCount+numFrames) — numFrames*spp samples
 * per pixel accumulated with the reference's progressive blend (Test.cpp:272-276,293-295) — over the rows
 * y_i = row0 + i*rowStep, i in [0,numRows) (TraceRowJob's [start,end) generalised for multi-GPU sharding).
 *   backbuffer        width*height*4 floats, row 0 = bottom, RGBA; read as `prev` and updated in place exactly like
 *                     the reference's (Test.cpp:293-296; alpha is preserved). With `packed` != 0 the buffer holds
 *                     only the numRows rendered rows, back to back.
 *   bufferOnDevice    0: host pointer (copied in/out inside the call; the call is synchronous like DrawTest)
 *                     1: device pointer (work is enqueued on `cudaStream`, no host synchronisation unless a ray
 *                        count is requested)
 *   outRayCount       NULL or receives the number of rays (every HitWorld call: camera, bounce, shadow —
 *                     Test.cpp:122,199) of this call; 64-bit because 3840x2160x64 spp exceeds INT_MAX.
 *   outRaysPerFrame   NULL or numFrames entries (exact mode only; fast mode fills entry 0 with the total).
 *   cudaStream        a cudaStream_t (NULL = the context's own stream).
 * numRows == 0 is an empty shard: nothing is traced or copied, the ray counts are 0.
 * Alpha: never written in exact mode (Maths.h:38). Fast mode keeps it too, except for a HOST buffer drawn with zero
 * `prev` weight (flags without kFlagProgressive, or frame 0): `prev` is then not uploaded and the rendered rows get
 * alpha = 0 — what every reference shell's zero-initialised buffer holds (TestWin.cpp:73-74).
 * Ordering: a context owns one scene and one set of counters; its draws execute in issue order even when they are
 * enqueued on different streams (each draw waits on the previous draw's end event). Use one context per thread. */
int tpt_draw(tpt_context* ctx, int frameCount, int numFrames, int width, int height,
             int row0, int numRows, int rowStep, int packed,
             float* backbuffer, int bufferOnDevice, unsigned testFlags, int mode,
             long long* outRayCount, long long* outRaysPerFrame, void* cudaStream);

/* Rays traced by all draws since the last call of this function (synchronises the stream). */
int tpt_read_ray_count(tpt_context* ctx, void* cudaStream, long long* outRays);
/* Device time (CUDA events on the launching stream) of the kernels of the most recent tpt_draw, in ms. */
int tpt_last_kernel_ms(tpt_context* ctx, float* outMs);
/* Number of kernel launches issued by the most recent tpt_draw. */
int tpt_last_launch_count(tpt_context* ctx);
/* Bytes the most recent tpt_set_scene copied host -> device: the packed scene blob, or 0 when its bytes were already
 * resident (a shell calls UpdateTest every frame; option "scene_upload_always" = 1 copies regardless). */
long long tpt_last_scene_upload_bytes(tpt_context* ctx);

/* Optional epilogue ("next" row of SURVEY §8f): linear float RGBA -> 8-bit sRGB RGBA with Y flip, the
 * presentation step of the reference shells (Cpp/Windows/PixelShader.hlsl:1-15). dst = width*height*4 bytes. */
int tpt_tonemap_srgb8(tpt_context* ctx, const float* image, int imageOnDevice, int width, int height,
                      unsigned char* dst, int dstOnDevice, void* cudaStream);
/* The reference's other two 8-bit conversions in the same pass: transfer 0 = the above, 1 = min(sqrtf(x)*255, 255), the
 * WebAssembly shell's cheap gamma (Cpp/Emscripten/main.cpp:67-79, which also flips Y), 2 = the C# TGA writer's
 * LinearToSRGB with its 255.9 truncation (Cs/Program.cs:34-68: bgr = 1, flipY = 0 — TGA rows are bottom-up like the
 * backbuffer's). */
int tpt_tonemap_rgba8(tpt_context* ctx, const float* image, int imageOnDevice, int width, int height,
                      unsigned char* dst, int dstOnDevice, int transfer, int bgr, int flipY, void* cudaStream);

/* Multi-GPU, one process per GPU: device memory that another process's kernels can write into directly over
 * NVLink/NVSwitch (CUDA IPC). The root rank allocates the image and exports a 64-byte handle; every other rank
 * opens it and passes the returned pointer as a device `backbuffer` to tpt_draw with its own row shard
 * (row0 = rank, rowStep = world, packed = 0): the trace kernel's 128-bit stores / vector reductions then land in
 * the root's HBM as the pixels finish — no separate gather. The reference has no counterpart (single device). */
int tpt_mem_alloc(tpt_context* ctx, unsigned long long bytes, void** outDevPtr);      /* zero-filled */
int tpt_mem_free(tpt_context* ctx, void* devPtr);
int tpt_mem_copy(tpt_context* ctx, void* dst, const void* src, unsigned long long bytes, int kind /*1 H2D, 2 D2H, 3 D2D*/);
int tpt_ipc_export(tpt_context* ctx, void* devPtr, void* outHandle64);
int tpt_ipc_open(tpt_context* ctx, const void* handle64, void** outDevPtr);
int tpt_ipc_close(tpt_context* ctx, void* devPtr);

/* Diagnostic used by the parity tests: evaluates the device-side libm restatement the exact mode uses
 * (toypathtracer_b200/csrc/tpt_libm.cuh) on n host floats. fn: 0 = sinf, 1 = cosf, 2 = powf(x, 5), 3 = powf(x, 1.0f/3.0f)
 * (the REFGPU mode's pow(x, 1.0/3.0), ComputeShader.hlsl:33). */
int tpt_debug_libm(tpt_context* ctx, int fn, const float* in, float* out, long long n);
/* Diagnostic used by the parity tests: nearest hit (HitWorld, Test.cpp:85-106) of n host rays {o.xyz, d.xyz} in the
 * current scene with one sweep form of the fast kernels (kform 0 reference form, 1 expanded, 2 expanded on packed pairs,
 * 3 conservative packed pass + reference-form decisions), tMin/tMax as Trace uses them. outId -1 = miss. */
int tpt_debug_hit(tpt_context* ctx, int kform, const float* rays6, int* outId, float* outT, long long n);
/* Diagnostic: device timestamps (ms since the start of the last progress-mode host draw) of the trace kernel's end,
 * each band copy's end and the draw's end. Returns minus the number of entries written. */
int tpt_debug_timeline(tpt_context* ctx, float* outMs, int capacity);

#ifdef __cplusplus
}
#endif
#endif
