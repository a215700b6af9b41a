// Host-side packing of the reference's raw scene export (20 B spheres, 36 B materials, emissive id list)
// into the device blob described in tpt_types.h. Mirrors what UpdateTest does on the CPU path
// (Cpp/Source/Test.cpp:321-339: invRadius, SoA fill with r^2, emissive list) and SpheresSoA's padding to a
// multiple of 4 with "impossible" spheres (Cpp/Source/Maths.h:370-388).
#pragma once
#include "tpt_types.h"
#include <vector>
#include <string.h>

namespace tpt {

inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

inline SceneBlobLayout scene_blob_layout(int count, int nLights)
{
    int simdCount = (count + 3) / 4 * 4;
    SceneBlobLayout L;
    uint32_t off = 0;
    L.offSph = off;        off += align16((uint32_t)simdCount * 16u);
    L.offInvRadius = off;  off += align16((uint32_t)simdCount * 4u);
    L.offLights = off;     off += align16((uint32_t)(nLights > 0 ? nLights : 1) * (uint32_t)sizeof(LightRec));
    L.geomBytes = off;
    L.offMatA = off;       off += align16((uint32_t)(count + 1) * 16u);
    L.offMatB = off;       off += align16((uint32_t)(count + 1) * 16u);
    L.offMatRi = off;      off += align16((uint32_t)(count + 1) * 4u);
    L.totalBytes = off;
    L.flags = 0;
    return L;
}

// emissives == nullptr: derive the list like UpdateTest (Test.cpp:333-338).
inline void pack_scene_blob(const Sphere20* spheres, const Material36* mats, int count,
                            const int* emissives, int emissiveCount,
                            std::vector<unsigned char>& blob, SceneBlobLayout& L, int& nLights, uint32_t sceneFlags = 0)
{
    std::vector<int> em;
    if (emissives) em.assign(emissives, emissives + emissiveCount);
    else
        for (int i = 0; i < count; ++i)
            if (mats[i].emissive[0] > 0 || mats[i].emissive[1] > 0 || mats[i].emissive[2] > 0) em.push_back(i);
    nLights = (int)em.size();
    L = scene_blob_layout(count, nLights);
    L.flags = sceneFlags;
    int simdCount = (count + 3) / 4 * 4;
    blob.assign(L.totalBytes, 0);
    Q4* sph = (Q4*)(blob.data() + L.offSph);
    float* invR = (float*)(blob.data() + L.offInvRadius);
    LightRec* lights = (LightRec*)(blob.data() + L.offLights);
    Q4* matA = (Q4*)(blob.data() + L.offMatA);
    Q4* matB = (Q4*)(blob.data() + L.offMatB);
    float* matRi = (float*)(blob.data() + L.offMatRi);
    for (int i = 0; i < simdCount; ++i)
    {
        if (i < count)
        {
            const Sphere20& s = spheres[i];
            sph[i].x = s.center[0]; sph[i].y = s.center[1]; sph[i].z = s.center[2];
            sph[i].w = s.radius * s.radius;          // Test.cpp:329
            invR[i] = 1.0f / s.radius;               // Maths.h:359
        }
        else
        {
            sph[i].x = sph[i].y = sph[i].z = 10000.0f; sph[i].w = 0.0f; invR[i] = 0.0f; // Maths.h:382-387
        }
    }
    for (int i = 0; i < count; ++i)
    {
        const Material36& m = mats[i];
        matA[i].x = m.albedo[0]; matA[i].y = m.albedo[1]; matA[i].z = m.albedo[2];
        memcpy(&matA[i].w, &m.type, 4);
        matB[i].x = m.emissive[0]; matB[i].y = m.emissive[1]; matB[i].z = m.emissive[2];
        matB[i].w = ((sceneFlags & kSceneMitsuba) && m.type == kMetal) ? 0.0f : m.roughness;   // Test.cpp:143-145
        matRi[i] = m.ri;
    }
    // entry [count]: all-zero (Lambert, black) — what the reference's out-of-bounds material read amounts to
    for (int j = 0; j < nLights; ++j)
    {
        int i = em[j];
        LightRec& R = lights[j];
        R.cx = spheres[i].center[0]; R.cy = spheres[i].center[1]; R.cz = spheres[i].center[2];
        R.radius = spheres[i].radius;
        R.ex = mats[i].emissive[0]; R.ey = mats[i].emissive[1]; R.ez = mats[i].emissive[2];
        R.id = i;
    }
}

inline SceneView scene_view_from_blob(const unsigned char* base, const SceneBlobLayout& L, int count, int nLights)
{
    SceneView v;
    v.sph = (const Q4*)(base + L.offSph);
    v.invRadius = (const float*)(base + L.offInvRadius);
    v.lights = (const LightRec*)(base + L.offLights);
    v.matA = (const Q4*)(base + L.offMatA);
    v.matB = (const Q4*)(base + L.offMatB);
    v.matRi = (const float*)(base + L.offMatRi);
    v.count = count;
    v.simdCount = (count + 3) / 4 * 4;
    v.nLights = nLights;
    v.flags = L.flags;
    v.sphShared = 0;
    return v;
}

} // namespace tpt
