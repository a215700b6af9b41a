// Data layouts shared by the host API and the kernels.
// The three raw structs are byte-for-byte the reference's scene export (GetSceneDesc, Cpp/Source/Test.cpp:377-384;
// sizes asserted by the reference's GPU shells at Cpp/Windows/TestWin.cpp:132-134):
//   Sphere   20 B  Cpp/Source/Maths.h:354-364     {float center[3]; float radius; float invRadius}
//   Material 36 B  Cpp/Source/Test.cpp:36-44      {int type; float albedo[3]; float emissive[3]; float roughness; float ri}
//   Camera   88 B  Cpp/Source/Maths.h:444-449     {origin, lowerLeftCorner, horizontal, vertical, uu, vv, ww (float3pack each); float lensRadius}
#pragma once
#include <stdint.h>

namespace tpt {

struct Sphere20 { float center[3]; float radius; float invRadius; };
struct Material36 { int type; float albedo[3]; float emissive[3]; float roughness; float ri; };
struct Camera88 { float origin[3], lowerLeftCorner[3], horizontal[3], vertical[3], uu[3], vv[3], ww[3]; float lensRadius; };

static_assert(sizeof(Sphere20) == 20, "Sphere layout");
static_assert(sizeof(Material36) == 36, "Material layout");
static_assert(sizeof(Camera88) == 88, "Camera layout");

enum MaterialType { kLambert = 0, kMetal = 1, kDielectric = 2 };

// 16-byte quad usable from host-only builds too (float4 on device).
struct alignas(16) Q4 { float x, y, z, w; };

// One record per emissive sphere (Test.cpp:321-338 builds the id list; Test.cpp:96-133 consumes it).
struct alignas(16) LightRec
{
    float cx, cy, cz, radius;   // s_Spheres[i]
    float ex, ey, ez;           // s_SphereMats[i].emissive
    int id;
};

// Device-side scene, SoA, padded like the reference's SpheresSoA (Maths.h:368-404):
//   sph[i]  = {centerX, centerY, centerZ, sqRadius}, i < simdCount, entries >= count are the
//             "impossible" spheres (centre 10000, r^2 = 0, invRadius 0)
//   matA[i] = {albedo.xyz, type (int bits)}, matB[i] = {emissive.xyz, roughness}, matC[i] = ri
//             i <= count; entry [count] is the all-zero material the reference's out-of-bounds read
//             amounts to (see DESIGN.md "padded-sphere hits").
struct SceneView
{
    const Q4* sph;
    const float* invRadius;
    const Q4* matA;
    const Q4* matB;
    const float* matRi;
    const LightRec* lights;
    int count, simdCount, nLights;
    uint32_t flags;       // kScene* bits
    uint32_t sphShared;   // device: 32-bit shared-memory address of sph[] (always staged), for ld.shared.v4
};

// Blob layout in global memory (one contiguous, 16 B-aligned allocation so a single bulk copy stages it
// into shared memory): [sph simdCount*16][matA (count+1)*16][matB (count+1)*16][lights nLights*32]
// [invRadius simdCount*4][matRi (count+1)*4], each section padded to 16 B.
struct SceneBlobLayout
{
    uint32_t offSph, offMatA, offMatB, offLights, offInvRadius, offMatRi, totalBytes;
    uint32_t geomBytes; // bytes of [sph .. end] needed by intersection + lights only (== totalBytes here)
    uint32_t flags;     // kScene* bits (travels to the kernels with the layout)
};

// DO_MITSUBA_COMPARE (Config.h:25) as a runtime switch: constant sky (Test.cpp:226-227) and zero Metal roughness
// (Test.cpp:143-145; applied when the scene is packed). Its third effect, zero aperture (Test.cpp:312-313), is camera
// data and therefore the caller's (the Test.h shim's UpdateTest builds the camera).
enum SceneFlags : uint32_t { kSceneMitsuba = 1u };

struct DrawParams
{
    Camera88 cam;
    int width, height;
    int row0, numRows;        // rows y_i = row0 + i*rowStep, i in [0,numRows)  (Test.cpp:266 TraceRowJob(start,end))
    int rowStep;              // 1 = contiguous band; world_size = rows interleaved across GPUs
    int packed;               // 0: row y_i lives at image row y_i (full image); 1: at image row i (packed band)
    int frame0, numFrames;    // frames [frame0, frame0+numFrames), N spp = N/spp reference frames
    int spp;                  // DO_SAMPLES_PER_PIXEL (Config.h:22)
    unsigned flags;           // kFlagAnimate = 1, kFlagProgressive = 2 (Test.h:4-8)
    int zeroAlpha;            // fast mode, prev weight 0: 1 = write alpha 0, 0 = keep the buffer's alpha (Maths.h:38: never written)
    float invWidth, invHeight;
    float* image;             // full image base, width*height*4 floats, row 0 = bottom (device)
    float* scratch;           // exact mode, numFrames > 1: [numFrames][numRows][width] float4 per-frame colours
    unsigned long long* rayCounter; // [numFrames] in exact mode, [1] in fast mode
    unsigned int* workCounter;      // persistent kernels: next tile
};

} // namespace tpt
