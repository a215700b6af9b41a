// Test-only: HOST build of the product's integrator source (toypathtracer_b200/csrc/tpt_integrator.cuh, EXACT
// instantiation, SerialHitter) so its logic can be checked bit-for-bit against the oracle on a machine without
// a GPU. Never linked into the product library; the product has no CPU path.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -mfma (fma only reaches the explicit __builtin_fma calls).
#include "../../toypathtracer_b200/csrc/tpt_integrator.cuh"
#include "../../toypathtracer_b200/csrc/tpt_scene_pack.h"
#include <thread>
#include <atomic>
#include <vector>

using namespace tpt;

extern "C" int sim_render_exact(const void* spheres, const void* mats, int count, const void* cam,
                                int w, int h, int frame0, int nframes, unsigned flags, int spp,
                                float* buf, long long* rays, int nthreads)
{
    std::vector<unsigned char> blob; SceneBlobLayout L; int nLights;
    pack_scene_blob((const Sphere20*)spheres, (const Material36*)mats, count, nullptr, 0, blob, L, nLights);
    SceneView sc = scene_view_from_blob(blob.data(), L, count, nLights);
    Camera88 c; memcpy(&c, cam, sizeof(c));
    float invW = 1.0f / w, invH = 1.0f / h;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    SerialHitter<true> hitter;
    for (int f = 0; f < nframes; ++f)
    {
        int frame = frame0 + f;
        float lerpFac = lerp_fac(frame, flags);
        std::atomic<int> next(0);
        std::atomic<long long> total(0);
        auto work = [&]() {
            long long mine = 0;
            for (;;)
            {
                int y = next.fetch_add(1);
                if (y >= h) break;
                uint32_t state = row_seed(y, frame);
                unsigned rc = 0;
                float* bb = buf + (size_t)y * w * 4;
                for (int x = 0; x < w; ++x, bb += 4)
                {
                    V3 col = pixel_exact(sc, c, x, y, spp, invW, invH, state, rc, hitter);
                    V3 prev = v3(bb[0], bb[1], bb[2]);
                    col = prev * lerpFac + col * (1.0f - lerpFac);
                    bb[0] = col.x; bb[1] = col.y; bb[2] = col.z;
                }
                mine += rc;
            }
            total += mine;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (rays) rays[f] = total.load();
    }
    return 0;
}


// Same, through the flat per-chain state machine (xchain_step) the batched LANES = 1 kernel uses.
extern "C" int sim_render_exact_flat(const void* spheres, const void* mats, int count, const void* cam,
                                     int w, int h, int frame0, int nframes, unsigned flags, int spp,
                                     float* buf, long long* rays, int nthreads)
{
    std::vector<unsigned char> blob; SceneBlobLayout L; int nLights;
    pack_scene_blob((const Sphere20*)spheres, (const Material36*)mats, count, nullptr, 0, blob, L, nLights);
    SceneView sc = scene_view_from_blob(blob.data(), L, count, nLights);
    Camera88 c; memcpy(&c, cam, sizeof(c));
    float invW = 1.0f / w, invH = 1.0f / h;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    SerialHitter<true> hitter;
    for (int f = 0; f < nframes; ++f)
    {
        int frame = frame0 + f;
        float lerpFac = lerp_fac(frame, flags);
        std::atomic<int> next(0);
        std::atomic<long long> total(0);
        auto work = [&]() {
            long long mine = 0;
            for (;;)
            {
                int y = next.fetch_add(1);
                if (y >= h) break;
                unsigned rc = 0;
                XChain ch;
                xchain_begin(ch, c, y, frame, invW, invH);
                float* row = buf + (size_t)y * w * 4;
                while (ch.x < w)
                {
                    V3 col;
                    const int x = ch.x;
                    if (xchain_step(sc, c, ch, y, spp, w, invW, invH, rc, hitter, col))
                    {
                        float* bb = row + (size_t)x * 4;
                        V3 prev = v3(bb[0], bb[1], bb[2]);
                        col = prev * lerpFac + col * (1.0f - lerpFac);
                        bb[0] = col.x; bb[1] = col.y; bb[2] = col.z;
                    }
                }
                mine += rc;
            }
            total += mine;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (rays) rays[f] = total.load();
    }
    return 0;
}


// Same, through the split form (xpath_sample -> event list -> xshade_event): the path stream of a whole row is produced
// FIRST (no colour arithmetic), the shade stream consumes it afterwards — the ordering freedom the two-warp kernel uses.
extern "C" int sim_render_exact_split(const void* spheres, const void* mats, int count, const void* cam,
                                      int w, int h, int frame0, int nframes, unsigned flags, int spp,
                                      float* buf, long long* rays, int nthreads)
{
    std::vector<unsigned char> blob; SceneBlobLayout L; int nLights;
    pack_scene_blob((const Sphere20*)spheres, (const Material36*)mats, count, nullptr, 0, blob, L, nLights);
    SceneView sc = scene_view_from_blob(blob.data(), L, count, nLights);
    Camera88 c; memcpy(&c, cam, sizeof(c));
    float invW = 1.0f / w, invH = 1.0f / h;
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    SerialHitter<true> hitter;
    struct Ev { int type, mid; V3 a, b, c; uint32_t rng; };
    for (int f = 0; f < nframes; ++f)
    {
        int frame = frame0 + f;
        float lerpFac = lerp_fac(frame, flags);
        std::atomic<int> next(0);
        std::atomic<long long> total(0);
        auto work = [&]() {
            long long mine = 0;
            std::vector<Ev> evs;
            for (;;)
            {
                int y = next.fetch_add(1);
                if (y >= h) break;
                uint32_t state = row_seed(y, frame);
                unsigned rc = 0;
                evs.clear();
                for (int x = 0; x < w; ++x)
                    for (int s = 0; s < spp; ++s)
                        xpath_sample(sc, c, x, y, invW, invH, state, rc, hitter,
                                     [&](int type, int mid, V3 a, V3 b, V3 cc, uint32_t rng) { evs.push_back(Ev{type, mid, a, b, cc, rng}); });
                size_t k = 0;
                float* bb = buf + (size_t)y * w * 4;
                for (int x = 0; x < w; ++x, bb += 4)
                {
                    V3 col = v3(0, 0, 0);
                    for (int s = 0; s < spp; ++s)
                    {
                        XShade sh; xshade_begin(sh);
                        V3 result;
                        for (;;)
                        {
                            const Ev& e = evs[k++];
                            if (xshade_event(sc, sh, e.type, e.mid, e.a, e.b, e.c, e.rng,
                                             [&](int mid, V3 pos, V3 normal, V3 rdir, V3 albedo, uint32_t rng) { return xlights_serial(sc, hitter, mid, pos, normal, rdir, albedo, rng); },
                                             result)) break;
                        }
                        col = col + result;
                    }
                    col = col * M<true>::div_(1.0f, (float)spp);
                    V3 prev = v3(bb[0], bb[1], bb[2]);
                    col = prev * lerpFac + col * (1.0f - lerpFac);
                    bb[0] = col.x; bb[1] = col.y; bb[2] = col.z;
                }
                mine += rc;
            }
            total += mine;
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (rays) rays[f] = total.load();
    }
    return 0;
}
