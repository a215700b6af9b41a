// TEST INFRASTRUCTURE — not product code.
// C-ABI harness around the UNMODIFIED reference renderer (Cpp/Source/Test.cpp, Maths.cpp, enkiTS/*),
// compiled where those sources lie under /root/reference by oracle/Makefile into oracle/_ref/libtoyref.so.
// Pattern follows the reference's own headless shells (Cpp/Emscripten/main.cpp:46-61, Cs/Program.cs:16-32):
//   InitializeTest(); per frame { UpdateTest(...); DrawTest(...); }  ShutdownTest();
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include "Test.h"
#include <string.h>
#include <chrono>

extern "C" {

// Renders frames [frame0, frame0+nframes) into `buf` (w*h*4 floats, caller-owned, NOT cleared here —
// the reference reads it as `prev` every frame, Test.cpp:293). rays[i] = outRayCount of frame i.
// seconds[i] (optional) = steady-clock time of UpdateTest+DrawTest for frame i (TestWin.cpp:310-321 style).
int ref_render(int w, int h, int frame0, int nframes, float time, unsigned flags,
               float* buf, long long* rays, double* seconds)
{
    static bool inited = false;
    if (!inited) { InitializeTest(); inited = true; }
    for (int i = 0; i < nframes; ++i)
    {
        int rc = 0;
        auto t0 = std::chrono::steady_clock::now();
        UpdateTest(time, frame0 + i, w, h, flags);
        DrawTest(time, frame0 + i, w, h, buf, rc, flags);
        auto t1 = std::chrono::steady_clock::now();
        if (rays) rays[i] = rc;
        if (seconds) seconds[i] = std::chrono::duration<double>(t1 - t0).count();
    }
    return 0;
}

// Raw scene export through the reference's own GetObjectCount/GetSceneDesc (Test.cpp:369-384),
// after an UpdateTest at the given size (Camera depends on aspect).
int ref_object_count(int* count, int* objSize, int* matSize, int* camSize)
{
    GetObjectCount(*count, *objSize, *matSize, *camSize);
    return 0;
}

int ref_scene_desc(float time, int frame, int w, int h, unsigned flags,
                   void* objects, void* materials, void* cam, int* emissives, int* emissiveCount)
{
    UpdateTest(time, frame, w, h, flags);
    GetSceneDesc(objects, materials, cam, emissives, emissiveCount);
    return 0;
}

void ref_shutdown() { /* scheduler lives for the process lifetime; enkiTS threads exit with it */ }

}
