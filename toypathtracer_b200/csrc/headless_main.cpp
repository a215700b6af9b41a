// Headless shell in the style of the reference's Cpp/Emscripten/main.cpp:46-61 and Cs/Program.cs:16-32: it only
// knows the six functions of Cpp/Source/Test.h and links against EITHER the reference's Test.cpp/Maths.cpp/enkiTS
// or this repo's libtoytest_b200.so — the same file proves the drop-in boundary both ways.
//   tpt_headless <width> <height> <frames> <flags> [out.tga]
// Prints per-frame ray counts, Mray/s (cumulative, like Program.cs:26-31) and an FNV-1a checksum of the float
// backbuffer bits; optionally writes a TGA (BGR, sqrt-gamma like Emscripten/main.cpp:67-79, no Y flip: TGA is
// bottom-up like the backbuffer).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

// Cpp/Source/Test.h:10-17
void InitializeTest();
void ShutdownTest();
void UpdateTest(float time, int frameCount, int screenWidth, int screenHeight, unsigned testFlags);
void DrawTest(float time, int frameCount, int screenWidth, int screenHeight, float* backbuffer, int& outRayCount, unsigned testFlags);

int main(int argc, char** argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 1280, h = argc > 2 ? atoi(argv[2]) : 720;
    const int frames = argc > 3 ? atoi(argv[3]) : 4;
    const unsigned flags = argc > 4 ? (unsigned)atoi(argv[4]) : 2u;
    std::vector<float> backbuffer((size_t)w * h * 4, 0.0f);
    InitializeTest();
    long long total = 0;
    double seconds = 0;
    for (int f = 0; f < frames; ++f)
    {
        int rays = 0;
        auto t0 = std::chrono::steady_clock::now();
        UpdateTest(0.0f, f, w, h, flags);
        DrawTest(0.0f, f, w, h, backbuffer.data(), rays, flags);
        auto t1 = std::chrono::steady_clock::now();
        if (f > 0) { total += rays; seconds += std::chrono::duration<double>(t1 - t0).count(); }   // frame 0 warms up
        printf("frame %d rays %d\n", f, rays);
    }
    uint64_t hsh = 1469598103934665603ull;
    const unsigned char* b = (const unsigned char*)backbuffer.data();
    for (size_t i = 0; i < backbuffer.size() * 4; ++i) { hsh ^= b[i]; hsh *= 1099511628211ull; }
    printf("checksum %016llx\n", (unsigned long long)hsh);
    if (seconds > 0) printf("%.1f Mray/s over %d frames\n", total / seconds * 1e-6, frames - 1);
    if (argc > 5)
    {
        FILE* fp = fopen(argv[5], "wb");
        if (fp)
        {
            unsigned char hdr[18] = {0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, (unsigned char)(w & 255), (unsigned char)(w >> 8),
                                     (unsigned char)(h & 255), (unsigned char)(h >> 8), 24, 0};
            fwrite(hdr, 1, 18, fp);
            std::vector<unsigned char> row((size_t)w * 3);
            for (int y = 0; y < h; ++y)
            {
                const float* p = backbuffer.data() + (size_t)y * w * 4;
                for (int x = 0; x < w; ++x, p += 4)
                    for (int c = 0; c < 3; ++c)
                        row[x * 3 + c] = (unsigned char)std::fmin(std::sqrt(std::fmax(p[2 - c], 0.0f)) * 255.0f, 255.0f);
                fwrite(row.data(), 1, row.size(), fp);
            }
            fclose(fp);
        }
    }
    ShutdownTest();
    return 0;
}
