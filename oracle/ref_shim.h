// TEST INFRASTRUCTURE ONLY -- nothing in the product path may include this file.
//
// The reference snapshot (/root/reference/projekt.cpp, projekt.h) is a fragment of a
// unity build: it has no #include and every math / platform symbol it uses is defined
// in files that are not part of the snapshot.  This header supplies exactly that missing
// layer so that the reference's scalar functions can be compiled *verbatim* (from where
// they lie under /root/reference, see oracle/Makefile) into oracle/_ref/libprojekt_ref.so.
//
// PARITY UNPINNED with respect to upstream: the semantics chosen here (SURVEY.md
// Appendix A) *define* the oracle, because upstream's own math layer is absent and the
// reference has no tests or golden vectors.  Every choice is stated next to the symbol
// together with the reference call sites (projekt.cpp:line) that constrain it.
#ifndef B200R_REF_SHIM_H
#define B200R_REF_SHIM_H

#include <stdint.h>
#include <string.h>
#include <math.h>
#include <immintrin.h>
#include <stdexcept>

typedef uint8_t u8;
typedef uint32_t u32;
typedef int32_t s32;
typedef float r32;
typedef int32_t b32;

#define internal static
#define global_variable static
#define BITMAP_BYTES_PER_PIXEL 4      // projekt.cpp:415
#define Pi32 3.14159265359f           // projekt.cpp:4143

// MergeSort(0,...) recurses with Half0 == 0 (projekt.cpp:20-33); Assert must be catchable.
struct ref_assert_failure : std::runtime_error {
    ref_assert_failure(const char *m) : std::runtime_error(m) {}
};
#define Assert(Expression) do { if(!(Expression)) throw ref_assert_failure(#Expression); } while(0)

// ---- vectors: member names used at projekt.cpp:84-89, 338-357, 4002-4008, 4096-4102 ----
union v2 { struct { r32 x, y; }; struct { r32 u, v; }; r32 E[2]; };
union v3 { struct { r32 x, y, z; }; struct { r32 r, g, b; }; struct { v2 xy; r32 Ignored0_; }; r32 E[3]; };
union v4 { struct { r32 x, y, z, w; }; struct { r32 r, g, b, a; }; r32 E[4]; };

inline v2 V2(r32 X, r32 Y) { v2 R; R.x = X; R.y = Y; return R; }
inline v2 V2i(s32 X, s32 Y) { v2 R; R.x = (r32)X; R.y = (r32)Y; return R; }      // projekt.cpp:430
inline v3 V3(r32 X, r32 Y, r32 Z) { v3 R; R.x = X; R.y = Y; R.z = Z; return R; }
inline v3 V3(v2 XY, r32 Z) { v3 R; R.x = XY.x; R.y = XY.y; R.z = Z; return R; }  // projekt.cpp:84
inline v4 V4(r32 X, r32 Y, r32 Z, r32 W) { v4 R; R.x = X; R.y = Y; R.z = Z; R.w = W; return R; }

inline v2 operator*(r32 A, v2 B) { return V2(A*B.x, A*B.y); }
inline v2 operator*(v2 B, r32 A) { return A*B; }
inline v2 &operator*=(v2 &B, r32 A) { B = A*B; return B; }                       // projekt.cpp:4006
inline v2 operator+(v2 A, v2 B) { return V2(A.x + B.x, A.y + B.y); }
inline v2 &operator+=(v2 &A, v2 B) { A = A + B; return A; }
inline v2 operator-(v2 A, v2 B) { return V2(A.x - B.x, A.y - B.y); }
inline v2 Hadamard(v2 A, v2 B) { return V2(A.x*B.x, A.y*B.y); }                  // projekt.cpp:432

inline v3 operator*(r32 A, v3 B) { return V3(A*B.x, A*B.y, A*B.z); }             // projekt.cpp:88-89
inline v3 operator*(v3 B, r32 A) { return A*B; }
inline v3 operator-(v3 A) { return V3(-A.x, -A.y, -A.z); }                       // projekt.cpp:475
inline v3 operator+(v3 A, v3 B) { return V3(A.x + B.x, A.y + B.y, A.z + B.z); }
inline v3 &operator+=(v3 &A, v3 B) { A = A + B; return A; }
inline v3 operator-(v3 A, v3 B) { return V3(A.x - B.x, A.y - B.y, A.z - B.z); }
// Inner: left-to-right sum of products (projekt.cpp:474, 3943, 4047)
inline r32 Inner(v3 A, v3 B) { return A.x*B.x + A.y*B.y + A.z*B.z; }
inline v3 Cross(v3 A, v3 B)                                                      // projekt.cpp:3943
{
    return V3(A.y*B.z - A.z*B.y, A.z*B.x - A.x*B.z, A.x*B.y - A.y*B.x);
}
// PIN: Normalize(a) = a * (1.0f / sqrtf(Inner(a,a)))  (Handmade lineage; projekt.cpp:471, 3926, 4029)
inline v3 Normalize(v3 A) { return A*(1.0f/sqrtf(Inner(A, A))); }

inline v4 operator*(r32 A, v4 B) { return V4(A*B.x, A*B.y, A*B.z, A*B.w); }      // projekt.cpp:4091
inline v4 operator+(v4 A, v4 B) { return V4(A.x + B.x, A.y + B.y, A.z + B.z, A.w + B.w); }
inline v4 &operator+=(v4 &A, v4 B) { A = A + B; return A; }                      // projekt.cpp:412, 548
inline v4 Hadamard(v4 A, v4 B) { return V4(A.x*B.x, A.y*B.y, A.z*B.z, A.w*B.w); } // projekt.cpp:466, 4042

inline r32 Clamp01(r32 V) { if(V < 0.0f) V = 0.0f; else if(V > 1.0f) V = 1.0f; return V; }  // projekt.cpp:474
inline v4 Clamp01(v4 V) { return V4(Clamp01(V.x), Clamp01(V.y), Clamp01(V.z), Clamp01(V.w)); } // :483

// PIN: cvtss2si under the default MXCSR (round-half-to-even; NaN/out-of-range -> 0x80000000).
// projekt.cpp:402-403, 433-434, 3988, 3999, 4093 (S32) and 490-493, 520-523 (U32).
inline s32 RoundR32ToS32(r32 V) { return _mm_cvtss_si32(_mm_set_ss(V)); }
inline u32 RoundR32ToU32(r32 V) { return (u32)_mm_cvtss_si32(_mm_set_ss(V)); }

inline r32 Maximum(r32 A, r32 B) { return (A > B) ? A : B; }                     // projekt.cpp:3999
inline r32 Sin(r32 A) { return sinf(A); }                                        // projekt.cpp:4168
inline r32 Cos(r32 A) { return cosf(A); }                                        // projekt.cpp:4164
inline void Copy(size_t Size, void *Source, void *Dest) { memcpy(Dest, Source, Size); } // projekt.cpp:2329

// ---- platform / render structs that projekt.h uses but does not define ----
struct loaded_bitmap                  // projekt.cpp:387, 414-416, 430-438
{
    s32 Width;
    s32 Height;
    s32 Pitch;
    void *Memory;
};
struct projective_transform           // projekt.cpp:79-89, 152-155
{
    r32 MetersToPixels;
    v2 ScreenCenter;
    r32 FocalLength;
    r32 DistanceAboveTarget;
};
struct light_info { v3 P; v4 Intensity; };        // projekt.cpp:469, 480, 4026-4027
struct light_data                                  // projekt.cpp:452-453, 466, 3892, 4010, 4023
{
    v4 AmbientIntensity;
    u32 LightCount;
    light_info *Lights;
};
struct game_render_commands           // projekt.cpp:170-171, 452, 458, 1017, 2325-2331, 4117
{
    u32 Width;                        // depth-buffer row stride in floats
    r32 *ZBuffer;
    u8 *ZMask;
    light_data LightData;
    projective_transform Transform;
    void *ThreadMemory;
    u32 ThreadMemorySize;
    u32 ThreadMemorySizeUsed;
    void *SortMemory;
};
struct platform_work_queue;
#define PLATFORM_WORK_QUEUE_CALLBACK(name) void name(platform_work_queue *Queue, void *Data)

#endif
