/* TEST INFRASTRUCTURE ONLY -- see raster_oracle.h for scope, pinning and build flags.
 * Plain C restatement of the reference's scalar Gouraud path; every function cites the
 * reference lines it follows.  No code here is shared with the CUDA product. */
#include "raster_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <xmmintrin.h>

/* ---- math layer, pinned exactly as oracle/ref_shim.h pins it ---- */

/* RoundR32ToS32: cvtss2si, round-half-even, NaN/overflow -> INT_MIN (projekt.cpp:402, 3988) */
static int32_t round_s32(float V) { return _mm_cvtss_si32(_mm_set_ss(V)); }
static uint32_t round_u32(float V) { return (uint32_t)round_s32(V); }            /* :520-523 */
static float clamp01(float V) { if(V < 0.0f) V = 0.0f; else if(V > 1.0f) V = 1.0f; return V; }
static float inner3(const float A[3], const float B[3])                          /* left to right */
{
    return A[0]*B[0] + A[1]*B[1] + A[2]*B[2];
}
/* Normalize(a) = a * (1/sqrt(a.a))   (projekt.cpp:3926, 4029) */
static void normalize3(const float A[3], float Out[3])
{
    float S = 1.0f/sqrtf(inner3(A, A));
    Out[0] = S*A[0]; Out[1] = S*A[1]; Out[2] = S*A[2];
}

/* projekt.cpp:74-93 */
void orc_project_vertex(const float Cam[3], const orc_transform *T, float Out[3])
{
    float Dist = T->DistanceAboveTarget - Cam[2];                /* :81 */
    Out[0] = 0.0f; Out[1] = 0.0f; Out[2] = 0.0f;                 /* :77 */
    if(Dist > 0.2f)                                              /* :82, :86 */
    {
        float S = (1.0f/Dist)*T->FocalLength;                    /* :88, scalar*scalar first */
        float Px = S*Cam[0];
        float Py = S*Cam[1];
        Out[0] = T->ScreenCenterX + T->MetersToPixels*Px;        /* :89 */
        Out[1] = T->ScreenCenterY + T->MetersToPixels*Py;
        Out[2] = Dist + T->MetersToPixels*0.0f;
    }
}

/* Gouraud vertex colour, projekt.cpp:4022-4062 (non-bitmap branch).  The reference lights
 * each edge end separately; the value depends only on the vertex, so it is computed once. */
static void light_vertex(const float Cam[3], const float Nrm[3], const float Col[4],
                         const orc_scene *Scene, float Out[4])
{
    float C[4] = {0, 0, 0, 0};
    for(uint32_t L = 0; L < Scene->LightCount; ++L)
    {
        const orc_light *Light = Scene->Lights + L;
        float ToLight[3] = { Light->P[0] - Cam[0], Light->P[1] - Cam[1], Light->P[2] - Cam[2] };
        float Dir[3];
        normalize3(ToLight, Dir);                                /* :4029 */
        if(L == 0)                                               /* :4032-4044 */
        {
            for(int i = 0; i < 4; ++i) C[i] = Col[i]*Scene->Ambient[i];
        }
        float Dot = clamp01(inner3(Dir, Nrm));                   /* :4047 */
        for(int i = 0; i < 4; ++i)                               /* :4058 */
        {
            C[i] = clamp01(C[i] + Dot*(Col[i]*Light->Intensity[i]));
        }
    }
    for(int i = 0; i < 4; ++i) Out[i] = C[i];
}

/* projekt.cpp:2-72.  Count == 0 is undefined in the reference (infinite recursion guarded by
 * Assert, :20-33); here it is a no-op. */
void orc_merge_sort(uint32_t Count, orc_edge *First, orc_edge *Temp)
{
    if(Count <= 1) return;
    if(Count == 2)                                               /* :9-19 */
    {
        if(First[0].YMin > First[1].YMin)
        {
            orc_edge T = First[0]; First[0] = First[1]; First[1] = T;
        }
        return;
    }
    uint32_t Half0 = Count/2, Half1 = Count - Half0;             /* :22-23 */
    orc_merge_sort(Half0, First, Temp);
    orc_merge_sort(Half1, First + Half0, Temp);
    uint32_t A = 0, B = Half0;
    for(uint32_t Out = 0; Out < Count; ++Out)                    /* :39-59 */
    {
        if(A == Half0) Temp[Out] = First[B++];
        else if(B == Count) Temp[Out] = First[A++];
        else if(First[A].YMin < First[B].YMin) Temp[Out] = First[A++];
        else Temp[Out] = First[B++];                             /* ties: right half first */
    }
    memcpy(First, Temp, Count*sizeof(orc_edge));                 /* :65-70 */
}

/* projekt.cpp:3882-4121, PhongShading == 0, Object->Bitmap == 0. */
int32_t orc_fill_edge_table(const float *Pos, const float *Col, const float *Nrm,
                            uint32_t VertexCount, const float P[3], const orc_scene *Scene,
                            orc_edge *Edges, orc_edge *Temp)
{
    return orc_fill_edge_table_ex(Pos, Col, Nrm, VertexCount, P, Scene, 0, Edges, Temp);
}

int32_t orc_fill_edge_table_ex(const float *Pos, const float *Col, const float *Nrm,
                               uint32_t VertexCount, const float P[3], const orc_scene *Scene,
                               int32_t Phong, orc_edge *Edges, orc_edge *Temp)
{
    return orc_fill_edge_table_tex(Pos, Col, Nrm, 0, VertexCount, P, Scene, Phong, Edges, Temp);
}

/* projekt.cpp:3882-4121.  Phong selects :4012-4019 instead of :4020-4064; UV != 0 stands for
 * Object->Bitmap != 0 (:4034, :4050, :4078). */
int32_t orc_fill_edge_table_tex(const float *Pos, const float *Col, const float *Nrm, const float *UV,
                                uint32_t VertexCount, const float P[3], const orc_scene *Scene,
                                int32_t Phong, orc_edge *Edges, orc_edge *Temp)
{
    static const float White[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
    static const uint32_t Indices[3][2] = { {0, 1}, {1, 2}, {2, 0} };   /* :3936-3941 */
    if(Scene->LightCount == 0 && !Phong) return -1;
    uint32_t TriangleCount = VertexCount/3;                      /* :3886 */
    uint32_t Visible = 0;
    for(uint32_t Tri = 0; Tri < TriangleCount; ++Tri)
    {
        float Cam[3][3], Proj[3][3], Lit[3][4];
        for(int v = 0; v < 3; ++v)
        {
            for(int c = 0; c < 3; ++c) Cam[v][c] = Pos[9*(size_t)Tri + 3*v + c] + P[c];  /* :3900 */
            orc_project_vertex(Cam[v], &Scene->Transform, Proj[v]);                    /* :3907 */
        }
        /* back-face test, :3926-3927, :3943 with Eye = (0,0,-1) (:3888) */
        float D1[3] = { Proj[1][0] - Proj[0][0], Proj[1][1] - Proj[0][1], Proj[1][2] - Proj[0][2] };
        float D2[3] = { Proj[2][0] - Proj[0][0], Proj[2][1] - Proj[0][1], Proj[2][2] - Proj[0][2] };
        float N1[3], N2[3];
        normalize3(D1, N1);
        normalize3(D2, N2);
        float Cx = N1[1]*N2[2] - N1[2]*N2[1];
        float Cy = N1[2]*N2[0] - N1[0]*N2[2];
        float Cz = N1[0]*N2[1] - N1[1]*N2[0];
        float Facing = 0.0f*Cx + 0.0f*Cy + -1.0f*Cz;
        if(!(Facing > 0.0f)) continue;

        for(int v = 0; v < 3; ++v)
        {
            if(Phong)                                            /* :4014-4015: colours stay unlit */
            {
                for(int i = 0; i < 4; ++i) Lit[v][i] = Col[12*(size_t)Tri + 4*v + i];
            }
            else
            {
                /* :4034-4060: a textured object is lit as if every vertex were white */
                light_vertex(Cam[v], Nrm + 9*(size_t)Tri + 3*v, UV ? White : Col + 12*(size_t)Tri + 4*v, Scene, Lit[v]);
            }
        }

        for(uint32_t e = 0; e < 3; ++e)                          /* :3947 */
        {
            uint32_t MinI = Indices[e][0], MaxI = Indices[e][1];
            if(Proj[MinI][1] > Proj[MaxI][1]) { uint32_t T = MinI; MinI = MaxI; MaxI = T; }  /* :3957 */
            const float *MinV = Proj[MinI], *MaxV = Proj[MaxI];
            if(!(MaxV[1] > 0)) continue;                         /* :3968 */

            orc_edge *E = Edges + Visible;
            E->YMax = round_s32(MaxV[1]);                        /* :3988 */
            float ClippedY = 0, T = 0.0f;
            if(MinV[1] < 0.0f)                                   /* :3993-3997 */
            {
                ClippedY = -MinV[1];
                T = (-MinV[1])/(MaxV[1] - MinV[1]);
            }
            float RoundedMin = (float)round_s32(MinV[1]);
            E->YMin = (int32_t)((0.0f > RoundedMin) ? 0.0f : RoundedMin);   /* :3999 */
            E->XMin = MinV[0];                                   /* :4000 */
            E->ZMin = Cam[MinI][2];                              /* :4001 */
            float FirstUV[2] = { 0, 0 }, SecondUV[2] = { 0, 0 };
            E->UMin = E->VMin = E->OneOverZMin = 0.0f;
            E->UGradient = E->VGradient = E->OneOverZGradient = 0.0f;
            if(UV)
            {
                /* z of a PROJECTED vertex is DistanceAboveTarget - camera z (:81, :89) */
                const float *U0 = UV + 6*(size_t)Tri + 2*MinI, *U1 = UV + 6*(size_t)Tri + 2*MaxI;   /* :3982-3983 */
                E->UMin = U0[0]/MinV[2];                         /* :4002 a true division ... */
                E->VMin = U0[1]/MinV[2];                         /* :4003 */
                E->OneOverZMin = 1.0f/MinV[2];                   /* :4004 */
                float InvMax = 1.0f/MaxV[2], InvMin = 1.0f/MinV[2];
                SecondUV[0] = InvMax*U1[0]; SecondUV[1] = InvMax*U1[1];      /* :4006 ... the gradients */
                FirstUV[0] = InvMin*U0[0];  FirstUV[1] = InvMin*U0[1];       /* :4008 use u*(1/z) */
            }
            if(MinV[1] - MaxV[1] != 0)                           /* :4066 */
            {
                float YDiff = (float)E->YMax - (float)E->YMin;   /* :4070 */
                E->ZGradient = (Cam[MaxI][2] - Cam[MinI][2])/YDiff;          /* :4072 */
                E->Gradient = (MaxV[0] - MinV[0])/(MaxV[1] - MinV[1]);       /* :4073 */
                E->XMin += ClippedY*E->Gradient;                 /* :4075 */
                E->ZMin += ClippedY*E->ZGradient;                /* :4076 */
                if(UV)                                           /* :4078-4089 */
                {
                    E->UGradient = (SecondUV[0] - FirstUV[0])/YDiff;
                    E->VGradient = (SecondUV[1] - FirstUV[1])/YDiff;
                    E->UMin += ClippedY*E->UGradient;
                    E->VMin += ClippedY*E->VGradient;
                    E->OneOverZGradient = ((1.0f/MaxV[2]) - E->OneOverZMin)/YDiff;
                    E->OneOverZMin += ClippedY*E->OneOverZGradient;
                }
                for(int i = 0; i < 4; ++i)                       /* :4091 */
                {
                    E->MinColor[i] = (1.0f - T)*Lit[MinI][i] + T*Lit[MaxI][i];
                }
                E->Left = (E->YMin == round_s32(Proj[Indices[e][0]][1])) ? 1 : 0;     /* :4093 */
                for(int i = 0; i < 4; ++i)                       /* :4096-4102 */
                {
                    E->ColorGradient[i] = (Lit[MaxI][i] - E->MinColor[i])/YDiff;
                }
                for(int i = 0; i < 3; ++i)                       /* :4017-4018, :4104-4109 */
                {
                    /* the start normal is NOT advanced by the top clip (only colour is, :4091) */
                    E->MinNormal[i] = Phong ? Nrm[9*(size_t)Tri + 3*MinI + i] : 0.0f;
                    E->NormalGradient[i] = Phong ? (Nrm[9*(size_t)Tri + 3*MaxI + i] - E->MinNormal[i])/YDiff : 0.0f;
                }
                E->Triangle = (int32_t)Tri;
                ++Visible;                                       /* :4068 */
            }
        }
    }
    orc_merge_sort(Visible, Edges, Temp);                        /* :4117 */
    return (int32_t)Visible;
}

/* Running state of one edge while it is in the active list. */
typedef struct active_edge {
    float X, Z, C[4];
    float N[3];                                                  /* Phong only */
    float U, V, W;                                               /* textured only: u/z, v/z, 1/z */
    const orc_edge *E;
} active_edge;

/* projekt.cpp:212-216 / 229-233: does New sort strictly before Old? */
static int edge_before(const active_edge *New, const active_edge *Old)
{
    return New->X < Old->X ||
           (New->X == Old->X &&
            (New->E->Gradient < Old->E->Gradient ||
             (New->E->Gradient == Old->E->Gradient && New->E->Left < Old->E->Left)));
}

/* projekt.cpp:306-425 (span set-up) and 510-538 (Gouraud pixel loop). */
/* UnprojectVertex, projekt.cpp:147-160 */
static void unproject_vertex(float X, float Y, float Z, const orc_transform *T, float Out[3])
{
    float Dist = T->DistanceAboveTarget - Z;                     /* :152 */
    float Inv = 1.0f/T->MetersToPixels;
    float Ax = Inv*(X - T->ScreenCenterX), Ay = Inv*(Y - T->ScreenCenterY);   /* :154 */
    float S = Dist/T->FocalLength;                               /* :155 */
    Out[0] = S*Ax; Out[1] = S*Ay; Out[2] = Z;
}

/* Per-pixel Phong colour, projekt.cpp:452-483.  pow(x,16) is the C++ pow(float,int) overload:
 * double precision, then narrowed to r32 (:478). */
static void phong_shade(const float Color[4], const float Normal[3], float X, float Row, float Z,
                        const orc_scene *Scene, float Out[4])
{
    float Final[4] = {0, 0, 0, 0};                               /* :448 */
    float Pw[3];
    unproject_vertex(X, Row, Z, &Scene->Transform, Pw);          /* :455-458 */
    for(uint32_t Li = 0; Li < Scene->LightCount; ++Li)
    {
        const orc_light *Light = Scene->Lights + Li;
        if(Li == 0) for(int i = 0; i < 4; ++i) Final[i] = Color[i]*Scene->Ambient[i];   /* :466 */
        float ToLight[3] = { Light->P[0] - Pw[0], Light->P[1] - Pw[1], Light->P[2] - Pw[2] };
        float Ld[3], Vd[3], Hd[3];
        normalize3(ToLight, Ld);                                 /* :471 */
        float Cos = clamp01(inner3(Normal, Ld));                 /* :474 */
        float Neg[3] = { -Pw[0], -Pw[1], -Pw[2] };
        normalize3(Neg, Vd);                                     /* :475 */
        float Sum[3] = { Ld[0] + Vd[0], Ld[1] + Vd[1], Ld[2] + Vd[2] };
        normalize3(Sum, Hd);                                     /* :476 */
        float Term = clamp01(inner3(Normal, Hd));                /* :477 */
        Term = (float)pow((double)Term, 16.0);                   /* :478 */
        for(int i = 0; i < 4; ++i)                               /* :480 */
        {
            Final[i] = Final[i] + (Cos*(Color[i]*Light->Intensity[i]) + Term*(1.0f*Light->Intensity[i]));
        }
    }
    for(int i = 0; i < 4; ++i) Out[i] = clamp01(Final[i]);       /* :483 */
}

/* projekt.cpp:427-446: nearest texel at Round(uv*(dim-1)); r,g,b,a = channel/255. */
static void sample_texture(const orc_texture *Tex, float U, float V, float W, float Out[4], orc_stats *Stats)
{
    float Inv = 1.0f/W;                                          /* :429 */
    float Fu = Inv*U, Fv = Inv*V;
    float Tx = Fu*(float)(Tex->Width - 1), Ty = Fv*(float)(Tex->Height - 1);   /* :430-432 */
    int32_t X = round_s32(Tx), Y = round_s32(Ty);                /* :433-434 */
    int Clamped = 0;
    if(X < 0) { X = 0; Clamped = 1; } else if(X > Tex->Width - 1) { X = Tex->Width - 1; Clamped = 1; }
    if(Y < 0) { Y = 0; Clamped = 1; } else if(Y > Tex->Height - 1) { Y = Tex->Height - 1; Clamped = 1; }
    if(Clamped && Stats) Stats->TexelClamps += 1;
    uint32_t Texel = *(const uint32_t *)((const uint8_t *)Tex->Memory + (size_t)X*4 + (size_t)Y*Tex->Pitch);   /* :436-438 */
    Out[3] = (float)((Texel >> 24) & 0xFF)/255.0f;               /* :440-443 */
    Out[0] = (float)((Texel >> 16) & 0xFF)/255.0f;
    Out[1] = (float)((Texel >> 8) & 0xFF)/255.0f;
    Out[2] = (float)((Texel >> 0) & 0xFF)/255.0f;
}

/* Compatibility switches: the two rules in which the reference's AVX fillers differ
 * from its scalar path and that can be stated on top of the scalar arithmetic -- the right end of a span is
 * exclusive (projekt.cpp:782-794, the end-clip masks) and the depth test is >= (projekt.cpp:3205).  Test
 * infrastructure state, set before a render and not thread safe. */
static int g_compat = 0;
void orc_set_compat(int32_t Flags) { g_compat = Flags; }

static void orc_fill_span(const active_edge *L, const active_edge *R, int32_t Row,
                          int32_t PrimIndex, orc_target *T, orc_stats *Stats,
                          const orc_scene *Scene, int32_t Phong, const orc_texture *Tex)
{
    float XDiff = roundf(R->X - L->X);                           /* :311-312 */
    float CInc[4], ZInc, NInc[3], UInc, VInc, WInc;
    if(XDiff != 0.0f)                                            /* :333-363 */
    {
        WInc = (R->W - L->W)/XDiff;                              /* :336 */
        UInc = (R->U - L->U)/XDiff; VInc = (R->V - L->V)/XDiff;  /* :338-342 */
        for(int i = 0; i < 4; ++i) CInc[i] = (R->C[i] - L->C[i])/XDiff;
        for(int i = 0; i < 3; ++i) NInc[i] = (R->N[i] - L->N[i])/XDiff;      /* :344-349 */
        ZInc = (R->Z - L->Z)/XDiff;
    }
    else
    {
        for(int i = 0; i < 4; ++i) CInc[i] = 0.0f;
        for(int i = 0; i < 3; ++i) NInc[i] = 0.0f;
        ZInc = 0.0f; UInc = VInc = WInc = 0.0f;
    }
    float U = L->U, V = L->V, W = L->W;                          /* :376-377 */
    float N[3] = { L->N[0], L->N[1], L->N[2] };                  /* :378 */
    float Z = L->Z;                                              /* :375 */
    float C[4] = { L->C[0], L->C[1], L->C[2], L->C[3] };         /* :379 */
    float XOffset = 0.0f;
    float LeftX = L->X;                                          /* :381-390 */
    if(LeftX < 0) { XOffset = -LeftX; LeftX = 0; }
    else if(LeftX >= (float)T->Width) { LeftX = (float)T->Width - 1; }
    float RightX = R->X;                                         /* :392-400 */
    if(RightX < 0) { RightX = 0; }
    else if(RightX >= (float)T->Width) { RightX = (float)T->Width - 1; }
    int32_t MinX = (int32_t)(float)round_s32(LeftX);             /* :402-406 */
    int32_t MaxX = (int32_t)(float)round_s32(RightX);
    if(g_compat & ORC_COMPAT_RIGHT_END_EXCLUSIVE) MaxX -= 1;     /* AVX fillers: [MinX, MaxX) */
    Z += XOffset*ZInc;                                           /* :408 */
    W += XOffset*WInc;                                           /* :409 */
    U += XOffset*UInc; V += XOffset*VInc;                        /* :410 */
    for(int i = 0; i < 3; ++i) N[i] += XOffset*NInc[i];          /* :411 */
    for(int i = 0; i < 4; ++i) C[i] += XOffset*CInc[i];          /* :412 */

    uint32_t *Pixel = (uint32_t *)((uint8_t *)T->Color + (size_t)MinX*4 + (size_t)Row*T->Pitch);  /* :414 */
    float *ZPixel = T->Z + MinX + (size_t)Row*T->ZStride;        /* :418 */
    int32_t *Prim = T->Prim ? T->Prim + MinX + (size_t)Row*T->ZStride : 0;
    if(Stats) Stats->SpanRows += 1;
    /* Column == Width is reachable: an end in [Width-0.5, Width) is not clamped (:387, :397) and
     * rounds up (:402-403).  The reference then writes through Row*Pitch + Width*4, i.e. into
     * column 0 of the NEXT row when rows are contiguous (Pitch == Width*4, ZBufferWidth == Width),
     * into row padding otherwise, and past the end of the buffer on the last row (undefined).
     * Defined here: the next-row write is kept (it is what the reference's image shows), the
     * padding / out-of-buffer writes are dropped. */
    const int Contiguous = (T->Pitch == T->Width*4) && ((int32_t)T->ZStride == T->Width);
    for(int32_t X = MinX; X <= MaxX; ++X)                        /* :423 */
    {
        if(X >= T->Width && !(Contiguous && Row + 1 < T->Height))
        {
            ++ZPixel; ++Pixel; if(Prim) ++Prim;
            if(Phong) { float Tn[3] = { N[0] + NInc[0], N[1] + NInc[1], N[2] + NInc[2] }; normalize3(Tn, N); }
            for(int i = 0; i < 4; ++i) C[i] = C[i] + CInc[i];
            Z += ZInc; U += UInc; V += VInc; W += WInc;
            continue;
        }
        if(Tex) sample_texture(Tex, U, V, W, C, Stats);          /* :427-446: replaces the running colour */
        float F[4] = { C[0], C[1], C[2], C[3] };                 /* :513 */
        if(Phong) phong_shade(C, N, (float)X, (float)Row, Z, Scene, F);     /* :450-483 */
        /* colour is r,g,b,a = C[0..3]; packed A R G B (:520-523) */
        uint32_t Color32 = (round_u32(F[3]*255.0f) << 24) | (round_u32(F[0]*255.0f) << 16) |
                           (round_u32(F[1]*255.0f) << 8) | (round_u32(F[2]*255.0f) << 0);
        if(Stats) Stats->Fragments += 1;
        if((g_compat & ORC_COMPAT_DEPTH_GE) ? (Z >= *ZPixel) : (Z > *ZPixel))     /* :525 (AVX single-thread variant: >=, :3205) */
        {
            *ZPixel = Z;
            *Pixel = Color32;
            if(Prim) *Prim = PrimIndex;
            if(Stats) Stats->DepthPasses += 1;
        }
        ++ZPixel; ++Pixel; if(Prim) ++Prim;
        if(Phong) { float Tn[3] = { N[0] + NInc[0], N[1] + NInc[1], N[2] + NInc[2] }; normalize3(Tn, N); }   /* :504 */
        for(int i = 0; i < 4; ++i) C[i] = C[i] + CInc[i];        /* :534 / :505 */
        Z += ZInc;                                               /* :535 / :506 */
        U += UInc; V += VInc; W += WInc;                         /* :507-508 / :536-537 */
    }
}

/* Level-1 active-edge walk of one triangle (projekt.cpp:173-303 and 542-597 restated for a
 * list of at most three edges with array storage instead of the intrusive linked list).
 * Defined behaviour where the reference is undefined:
 *   - a row whose list holds fewer than two edges draws nothing and steps nothing
 *     (the reference reads ListHead->YMax / ->Next through a null pointer, :262, :301);
 *   - after two edges cross, they are exchanged in the list (the reference exchanges them
 *     but leaves ListHead/ListTail stale, :562-572, and walks off the list one row later,
 *     :274-278): bit 2 of the result reports that case. */
int32_t orc_draw_triangle(const orc_edge *Edges, uint32_t EdgeCount, int32_t PrimIndex,
                          orc_target *T, orc_stats *Stats)
{
    return orc_draw_triangle_ex(Edges, EdgeCount, PrimIndex, T, Stats, 0, 0);
}

int32_t orc_draw_triangle_ex(const orc_edge *Edges, uint32_t EdgeCount, int32_t PrimIndex,
                             orc_target *T, orc_stats *Stats, const orc_scene *Scene, int32_t Phong)
{
    return orc_draw_triangle_tex(Edges, EdgeCount, PrimIndex, T, Stats, Scene, Phong, 0);
}

int32_t orc_draw_triangle_tex(const orc_edge *Edges, uint32_t EdgeCount, int32_t PrimIndex,
                              orc_target *T, orc_stats *Stats, const orc_scene *Scene, int32_t Phong,
                              const orc_texture *Tex)
{
    if(EdgeCount == 0) return 0;
    if(EdgeCount > 3) EdgeCount = 3;
    int32_t FirstRow = Edges[0].YMin;                            /* :173 */
    int32_t MaxRow = Edges[0].YMax;                              /* :176-185 */
    for(uint32_t e = 1; e < EdgeCount; ++e) if(MaxRow < Edges[e].YMax) MaxRow = Edges[e].YMax;
    int32_t MaxY = MaxRow;                                       /* :187-196 */
    if(MaxY > T->Height) MaxY = T->Height;

    active_edge List[3];
    uint32_t Count = 0;
    int32_t Result = 0;
    for(int32_t Row = FirstRow; Row < MaxY; ++Row)               /* :198 */
    {
        for(uint32_t e = 0; e < EdgeCount; ++e)                  /* :202-260 insertion */
        {
            if(Edges[e].YMin != Row) continue;
            active_edge New;
            New.X = Edges[e].XMin; New.Z = Edges[e].ZMin; New.E = Edges + e;
            for(int i = 0; i < 4; ++i) New.C[i] = Edges[e].MinColor[i];
            for(int i = 0; i < 3; ++i) New.N[i] = Edges[e].MinNormal[i];
            New.U = Edges[e].UMin; New.V = Edges[e].VMin; New.W = Edges[e].OneOverZMin;
            uint32_t At = Count;
            for(uint32_t k = 0; k < Count; ++k)
            {
                if(edge_before(&New, &List[k])) { At = k; break; }
            }
            for(uint32_t k = Count; k > At; --k) List[k] = List[k - 1];
            List[At] = New;
            ++Count;
        }
        uint32_t Kept = 0;                                       /* :262-296 expiry */
        for(uint32_t k = 0; k < Count; ++k)
        {
            if(List[k].E->YMax <= Row) continue;
            List[Kept++] = List[k];
        }
        Count = Kept;
        if(Count < 2) continue;                                  /* defined: nothing happens */

        active_edge *L = &List[0], *R = &List[1];                /* :300-303, first pair only */
        orc_fill_span(L, R, Row, PrimIndex, T, Stats, Scene, Phong, Tex);   /* Row >= 0 always (:308) */
        Result |= 1;
        L->X += L->E->Gradient;      R->X += R->E->Gradient;     /* :542-543 */
        L->Z += L->E->ZGradient;     R->Z += R->E->ZGradient;    /* :545-546 */
        for(int i = 0; i < 4; ++i)                               /* :548-549 */
        {
            L->C[i] += L->E->ColorGradient[i];
            R->C[i] += R->E->ColorGradient[i];
        }
        if(Phong)                                                /* :551-552 */
        {
            float Tl[3], Tr[3];
            for(int i = 0; i < 3; ++i) { Tl[i] = L->N[i] + L->E->NormalGradient[i]; Tr[i] = R->N[i] + R->E->NormalGradient[i]; }
            normalize3(Tl, L->N);
            normalize3(Tr, R->N);
        }
        L->U += L->E->UGradient; L->V += L->E->VGradient; L->W += L->E->OneOverZGradient;   /* :554-556 */
        R->U += R->E->UGradient; R->V += R->E->VGradient; R->W += R->E->OneOverZGradient;   /* :558-560 */
        if(L->X > R->X)                                          /* :562-572 */
        {
            active_edge Tmp = *L; *L = *R; *R = Tmp;
            if(Row + 1 < MaxY) Result |= 2;
        }
    }
    return Result;
}

/* ---- level 0: the whole-object list, link by link (see raster_oracle.h) ---------------------- */
typedef struct object_walk {
    active_edge *A;            /* running values; the reference keeps them in the edge records */
    int32_t *Next;             /* edge_info::Next as an index, -1 = null */
    uint32_t Count;
} object_walk;

/* projekt.cpp:212-216 / 229-233 on list members */
static int walk_before(const object_walk *W, int32_t New, int32_t Old)
{
    return edge_before(&W->A[New], &W->A[Old]);
}

int32_t orc_draw_object(const orc_edge *Edges, uint32_t EdgeCount, int32_t PrimBase,
                        orc_target *T, orc_stats *Stats, const orc_scene *Scene, int32_t Phong,
                        const orc_texture *Tex, int32_t *SpansDrawn)
{
    if(SpansDrawn) *SpansDrawn = 0;
    if(EdgeCount == 0) return 0;
    int32_t FirstRow = Edges[0].YMin;                            /* :173 */
    int32_t MaxRow = Edges[0].YMax;                              /* :176-185 */
    for(uint32_t e = 1; e < EdgeCount; ++e) if(MaxRow < Edges[e].YMax) MaxRow = Edges[e].YMax;
    int32_t MaxY = MaxRow;                                       /* :187-196 */
    if(MaxY > T->Height) MaxY = T->Height;

    object_walk W;
    W.Count = EdgeCount;
    W.A = (active_edge *)malloc(sizeof(active_edge)*EdgeCount);
    W.Next = (int32_t *)malloc(sizeof(int32_t)*EdgeCount);
    for(uint32_t e = 0; e < EdgeCount; ++e)
    {
        active_edge *N = &W.A[e];
        N->X = Edges[e].XMin; N->Z = Edges[e].ZMin; N->E = Edges + e;
        for(int i = 0; i < 4; ++i) N->C[i] = Edges[e].MinColor[i];
        for(int i = 0; i < 3; ++i) N->N[i] = Edges[e].MinNormal[i];
        N->U = Edges[e].UMin; N->V = Edges[e].VMin; N->W = Edges[e].OneOverZMin;
        W.Next[e] = -1;                                          /* FillEdgeTable: Next = 0 (:4094) */
    }
    int32_t *Next = W.Next;
    active_edge *A = W.A;
    int32_t Head = -1, Tail = -1;                                /* :189-190 */
    int32_t Result = 0, Spans = 0;
    /* any loop of the reference that follows Next more often than this has entered a cycle */
    const uint64_t Fuse = 4ull*EdgeCount + 64;
#define NULL_DEREF() do { Result |= 2; goto done; } while(0)

    for(int32_t Row = FirstRow; Row < MaxY; ++Row)               /* :198 */
    {
        for(uint32_t e = 0; e < EdgeCount; ++e)                  /* :202-260 insertion */
        {
            if(Edges[e].YMin != Row) continue;
            const int32_t Cur = (int32_t)e;
            if(Head >= 0)
            {
                if(walk_before(&W, Cur, Head)) { Next[Cur] = Head; Head = Cur; }        /* :212-220 */
                else
                {
                    int32_t Compared = Head, Previous = Head;                            /* :223-224 */
                    uint64_t Steps = 0;
                    while(Compared != Tail)                                               /* :225 */
                    {
                        Compared = Next[Compared];                                        /* :227 */
                        if(Compared < 0 || ++Steps > Fuse) NULL_DEREF();                  /* stale ListTail */
                        if(walk_before(&W, Cur, Compared))                                /* :229-233 */
                        {
                            Next[Cur] = Compared; Next[Previous] = Cur; Compared = Tail;  /* :235-237 */
                        }
                        else Previous = Compared;                                         /* :241 */
                    }
                    if(Previous == Compared) { Next[Tail] = Cur; Tail = Cur; }            /* :246-250 */
                }
            }
            else { Head = Cur; Tail = Head; }                                             /* :254-258 */
        }

        for(uint64_t Steps = 0;; ++Steps)                        /* :262-267 expiry at the head */
        {
            if(Head < 0 || Steps > Fuse) NULL_DEREF();           /* the list ran empty */
            if(!(A[Head].E->YMax <= Row)) break;
            int32_t Removed = Head; Head = Next[Head]; Next[Removed] = -1;
        }
        {
            int32_t Previous = Head, Checked = Head;             /* :269-296 expiry behind the head */
            uint64_t Steps = 0;
            while(Checked != Tail)
            {
                Checked = Next[Checked];
                if(Checked < 0 || ++Steps > Fuse) NULL_DEREF();
                if(A[Checked].E->YMax <= Row)
                {
                    if(Checked == Tail) { Tail = Previous; Next[Tail] = -1; Checked = Tail; }
                    else { Next[Previous] = Next[Checked]; Checked = Previous; }
                }
                Previous = Checked;
            }
        }

        int32_t PrevCur = -1, PrevNext = -1;                     /* :298-303 */
        int32_t Cur = Head, Nxt = Next[Cur];                     /* Head >= 0 here */
        uint64_t Pairs = 0;
        while(Nxt >= 0)
        {
            if(++Pairs > Fuse) NULL_DEREF();
            orc_fill_span(&A[Cur], &A[Nxt], Row, PrimBase + Spans, T, Stats, Scene, Phong, Tex);   /* :306-540 */
            ++Spans; Result |= 1;
            active_edge *L = &A[Cur], *R = &A[Nxt];
            L->X += L->E->Gradient;      R->X += R->E->Gradient;     /* :542-543 */
            L->Z += L->E->ZGradient;     R->Z += R->E->ZGradient;    /* :545-546 */
            for(int i = 0; i < 4; ++i) { L->C[i] += L->E->ColorGradient[i]; R->C[i] += R->E->ColorGradient[i]; }   /* :548-549 */
            if(Phong)                                                /* :551-552 */
            {
                float Tl[3], Tr[3];
                for(int i = 0; i < 3; ++i) { Tl[i] = L->N[i] + L->E->NormalGradient[i]; Tr[i] = R->N[i] + R->E->NormalGradient[i]; }
                normalize3(Tl, L->N);
                normalize3(Tr, R->N);
            }
            L->U += L->E->UGradient; L->V += L->E->VGradient; L->W += L->E->OneOverZGradient;   /* :554-556 */
            R->U += R->E->UGradient; R->V += R->E->VGradient; R->W += R->E->OneOverZGradient;   /* :558-560 */

            if(A[Cur].X > A[Nxt].X)                                  /* :562-572: exchange inside the pair */
            {
                Next[Cur] = Next[Nxt];
                Next[Nxt] = Cur;
                if(PrevNext >= 0) Next[PrevNext] = Nxt;
                Cur = Nxt;
                Nxt = Next[Cur];
            }
            if(PrevNext >= 0)                                        /* :574-584: exchange across pairs */
            {
                if(A[PrevNext].X > A[Cur].X)
                {
                    Next[PrevNext] = Next[Cur];
                    Next[Cur] = PrevNext;
                    Next[PrevCur] = Cur;
                    PrevNext = Cur;
                    Cur = Next[PrevNext];
                    if(Cur < 0) NULL_DEREF();
                }
            }
            PrevCur = Cur; PrevNext = Nxt;                           /* :586-587 */
            if(Nxt < 0) NULL_DEREF();
            if(Next[Nxt] >= 0) { Cur = Next[Nxt]; Nxt = Next[Cur]; } /* :589-597 */
            else Nxt = -1;
        }
    }
done:
#undef NULL_DEREF
    free(W.A); free(W.Next);
    if(SpansDrawn) *SpansDrawn = Spans;
    return Result;
}

int32_t orc_render_object(const float *Pos, const float *Col, const float *Nrm, const float *UV,
                          uint32_t VertexCount, const float P[3], const orc_scene *Scene, int32_t Phong,
                          const orc_texture *Tex, orc_target *T, int32_t PrimBase, orc_stats *Stats)
{
    if(VertexCount < 3) return 0;
    orc_edge *Edges = (orc_edge *)malloc(sizeof(orc_edge)*VertexCount);
    orc_edge *Temp = (orc_edge *)malloc(sizeof(orc_edge)*VertexCount);
    const int Textured = UV && Tex;
    int32_t Count = orc_fill_edge_table_tex(Pos, Col, Nrm, Textured ? UV : 0, VertexCount, P, Scene, Phong, Edges, Temp);
    int32_t R = Count;
    if(Count > 0)
    {
        if(Stats) { Stats->Triangles += VertexCount/3; }
        R = orc_draw_object(Edges, (uint32_t)Count, PrimBase, T, Stats, Scene, Phong, Textured ? Tex : 0, 0);
        if(Stats && (R & 2)) Stats->RefWouldCrash += 1;
    }
    free(Edges); free(Temp);
    return R;
}

/* One triangle = one object: FillEdgeTable on the 3-vertex object, then the level-1 walk. */
static int32_t render_one_tex(const float *Pos, const float *Col, const float *Nrm, const float *UV, uint32_t Tri,
                              const float P[3], const orc_scene *Scene, int32_t Phong, const orc_texture *Tex,
                              orc_target *T, int32_t PrimIndex, orc_stats *Stats)
{
    orc_edge Edges[3], Temp[3];
    const int Textured = UV && Tex;
    int32_t Count = orc_fill_edge_table_tex(Pos + 9*(size_t)Tri, Col + 12*(size_t)Tri, Nrm + 9*(size_t)Tri,
                                            Textured ? UV + 6*(size_t)Tri : 0, 3, P, Scene, Phong, Edges, Temp);
    if(Count < 0) return Count;
    if(Stats) { Stats->Triangles += 1; if(Count > 0) Stats->Visible += 1; }
    int32_t R = orc_draw_triangle_tex(Edges, (uint32_t)Count, PrimIndex, T, Stats, Scene, Phong, Textured ? Tex : 0);
    if(Stats && (R & 2)) Stats->RefWouldCrash += 1;
    return R;
}

static int32_t render_one_ex(const float *Pos, const float *Col, const float *Nrm, uint32_t Tri,
                             const float P[3], const orc_scene *Scene, int32_t Phong, orc_target *T,
                             int32_t PrimIndex, orc_stats *Stats)
{
    return render_one_tex(Pos, Col, Nrm, 0, Tri, P, Scene, Phong, 0, T, PrimIndex, Stats);
}

static int32_t render_one(const float *Pos, const float *Col, const float *Nrm, uint32_t Tri,
                          const float P[3], const orc_scene *Scene, orc_target *T,
                          int32_t PrimIndex, orc_stats *Stats)
{
    return render_one_ex(Pos, Col, Nrm, Tri, P, Scene, 0, T, PrimIndex, Stats);
}

int32_t orc_render_triangles(const float *Pos, const float *Col, const float *Nrm,
                             uint32_t TriangleCount, const float P[3], const orc_scene *Scene,
                             orc_target *Target, int32_t PrimBase, uint8_t *WouldCrash,
                             orc_stats *Stats)
{
    return orc_render_triangles_ex(Pos, Col, Nrm, TriangleCount, P, Scene, 0, Target, PrimBase, WouldCrash, Stats);
}

int32_t orc_render_triangles_ex(const float *Pos, const float *Col, const float *Nrm,
                                uint32_t TriangleCount, const float P[3], const orc_scene *Scene,
                                int32_t Phong, orc_target *Target, int32_t PrimBase, uint8_t *WouldCrash,
                                orc_stats *Stats)
{
    if(Scene->LightCount == 0 && !Phong) return -1;
    for(uint32_t Tri = 0; Tri < TriangleCount; ++Tri)
    {
        int32_t R = render_one_ex(Pos, Col, Nrm, Tri, P, Scene, Phong, Target, PrimBase + (int32_t)Tri, Stats);
        if(WouldCrash) WouldCrash[Tri] = (uint8_t)((R & 2) ? 1 : 0);
    }
    return 0;
}

int32_t orc_render_triangles_tex(const float *Pos, const float *Col, const float *Nrm, const float *UV,
                                 uint32_t TriangleCount, const float P[3], const orc_scene *Scene,
                                 int32_t Phong, const orc_texture *Texture, orc_target *Target,
                                 int32_t PrimBase, uint8_t *WouldCrash, orc_stats *Stats)
{
    if(Scene->LightCount == 0 && !Phong) return -1;
    for(uint32_t Tri = 0; Tri < TriangleCount; ++Tri)
    {
        int32_t R = render_one_tex(Pos, Col, Nrm, UV, Tri, P, Scene, Phong, Texture, Target,
                                   PrimBase + (int32_t)Tri, Stats);
        if(WouldCrash) WouldCrash[Tri] = (uint8_t)((R & 2) ? 1 : 0);
    }
    return 0;
}

typedef struct worker_args {
    const float *Pos, *Col, *Nrm;
    uint32_t First, Last;
    const float *P;
    const orc_scene *Scene;
    orc_target Target;
    orc_stats Stats;
} worker_args;

static void *worker_main(void *Arg)
{
    worker_args *W = (worker_args *)Arg;
    for(uint32_t Tri = W->First; Tri < W->Last; ++Tri)
    {
        render_one(W->Pos, W->Col, W->Nrm, Tri, W->P, W->Scene, &W->Target, (int32_t)Tri, &W->Stats);
    }
    return 0;
}

/* Worker t renders the contiguous range t into a private copy of the target; copies are
 * folded in range order with the reference's depth rule (strict >, projekt.cpp:525), which
 * reproduces the single-thread result exactly (first submitted wins equal depth). */
int32_t orc_render_triangles_mt(const float *Pos, const float *Col, const float *Nrm,
                                uint32_t TriangleCount, const float P[3], const orc_scene *Scene,
                                orc_target *Target, uint32_t Threads, orc_stats *Stats)
{
    if(Scene->LightCount == 0) return -1;
    if(Threads < 1) Threads = 1;
    worker_args *Args = (worker_args *)calloc(Threads, sizeof(worker_args));
    pthread_t *Ids = (pthread_t *)calloc(Threads, sizeof(pthread_t));
    size_t ZCount = (size_t)Target->ZStride*(size_t)Target->Height;
    size_t CBytes = (size_t)Target->Pitch*(size_t)Target->Height;
    for(uint32_t t = 0; t < Threads; ++t)
    {
        worker_args *W = Args + t;
        W->Pos = Pos; W->Col = Col; W->Nrm = Nrm; W->P = P; W->Scene = Scene;
        W->First = (uint32_t)(((uint64_t)TriangleCount*t)/Threads);
        W->Last = (uint32_t)(((uint64_t)TriangleCount*(t + 1))/Threads);
        W->Target = *Target;
        if(t > 0)
        {
            W->Target.Color = (uint32_t *)malloc(CBytes);
            W->Target.Z = (float *)malloc(ZCount*sizeof(float));
            W->Target.Prim = Target->Prim ? (int32_t *)malloc(ZCount*sizeof(int32_t)) : 0;
            memcpy(W->Target.Color, Target->Color, CBytes);
            memcpy(W->Target.Z, Target->Z, ZCount*sizeof(float));
            if(Target->Prim) memcpy(W->Target.Prim, Target->Prim, ZCount*sizeof(int32_t));
        }
        pthread_create(&Ids[t], 0, worker_main, W);
    }
    for(uint32_t t = 0; t < Threads; ++t) pthread_join(Ids[t], 0);
    for(uint32_t t = 0; t < Threads; ++t)
    {
        worker_args *W = Args + t;
        if(Stats)
        {
            Stats->Triangles += W->Stats.Triangles;     Stats->Visible += W->Stats.Visible;
            Stats->SpanRows += W->Stats.SpanRows;       Stats->Fragments += W->Stats.Fragments;
            Stats->DepthPasses += W->Stats.DepthPasses; Stats->RefWouldCrash += W->Stats.RefWouldCrash;
        }
        if(t == 0) continue;
        for(int32_t Y = 0; Y < Target->Height; ++Y)
        {
            uint32_t *DstC = (uint32_t *)((uint8_t *)Target->Color + (size_t)Y*Target->Pitch);
            uint32_t *SrcC = (uint32_t *)((uint8_t *)W->Target.Color + (size_t)Y*Target->Pitch);
            float *DstZ = Target->Z + (size_t)Y*Target->ZStride;
            float *SrcZ = W->Target.Z + (size_t)Y*Target->ZStride;
            for(int32_t X = 0; X < Target->Width; ++X)
            {
                if(SrcZ[X] > DstZ[X])
                {
                    DstZ[X] = SrcZ[X]; DstC[X] = SrcC[X];
                    if(Target->Prim) Target->Prim[(size_t)Y*Target->ZStride + X] =
                        W->Target.Prim[(size_t)Y*Target->ZStride + X];
                }
            }
        }
        free(W->Target.Color); free(W->Target.Z); free(W->Target.Prim);
    }
    free(Args); free(Ids);
    return 0;
}

/* Layout mirrors of the two reference structs the fallback receives (oracle/ref_shim.h). */
typedef struct ref_loaded_bitmap { int32_t Width, Height, Pitch; void *Memory; } ref_loaded_bitmap;
typedef struct ref_commands_head { uint32_t Width; float *ZBuffer; } ref_commands_head;

void orc_ref_fallback(void *User, uint32_t TriangleIndex, void *RefLoadedBitmap,
                      void *RefGameRenderCommands)
{
    orc_fallback_ctx *Ctx = (orc_fallback_ctx *)User;
    ref_loaded_bitmap *B = (ref_loaded_bitmap *)RefLoadedBitmap;
    ref_commands_head *C = (ref_commands_head *)RefGameRenderCommands;
    orc_target T;
    T.Width = B->Width; T.Height = B->Height; T.Pitch = B->Pitch;
    T.Color = (uint32_t *)B->Memory; T.Z = C->ZBuffer; T.ZStride = C->Width; T.Prim = 0;
    render_one_tex(Ctx->Pos, Ctx->Col, Ctx->Nrm, Ctx->UV, TriangleIndex, Ctx->P, Ctx->Scene, Ctx->Phong,
                   Ctx->Texture, &T, (int32_t)TriangleIndex, 0);
}


/* ---- ConstructSphere, projekt.cpp:4123-4289, with StepCount as a parameter (the reference
 * hard-codes 24, :4129).  Same operations in the same order on binary32: Sin/Cos are libm's
 * sinf/cosf (ref_shim.h pins them the same way), the colour ramp is accumulated row by row
 * (:4286), pole UVs are (x, z) of the unit vertex (:4180, :4186), the others (x+1)/2, (y+1)/2.
 * Pinned at StepCount 24 against the verbatim function (tests/test_oracle_golden.py); config C5
 * uses StepCount 708.  Returns the vertex count: 3*(4*S*S - 4*S) for S >= 2. */
static void sphere_vertex(float Sx, float Sy, float Sz, float Radius, float U, float V, const float Color[4],
                          float *Pos, float *Col, float *Nrm, float *UV, uint32_t At)
{
    Pos[3*At + 0] = Radius*Sx; Pos[3*At + 1] = Radius*Sy; Pos[3*At + 2] = Radius*Sz;
    Nrm[3*At + 0] = Sx; Nrm[3*At + 1] = Sy; Nrm[3*At + 2] = Sz;
    UV[2*At + 0] = U; UV[2*At + 1] = V;
    for(int i = 0; i < 4; ++i) Col[4*At + i] = Color[i];
}

uint32_t orc_construct_sphere(uint32_t StepCount, float *Pos, float *Col, float *Nrm, float *UV)
{
    uint32_t N = 0;
    const float Radius = 0.5f;
    const float Pi32 = 3.14159265359f;
    const float Up[4] = {1.0f, 0.0f, 0.0f, 1.0f}, Down[4] = {0.0f, 1.0f, 0.0f, 1.0f};
    float Inc[4], Cur[4];
    for(int i = 0; i < 4; ++i) { Inc[i] = (Down[i] - Up[i])/(float)StepCount; Cur[i] = Up[i]; }   /* :4135-4141 */
    const float InclInc = Pi32/StepCount;                                   /* :4143 */
    const float AzimInc = (2.0f*Pi32)/(StepCount*2);                        /* :4144 */
    for(uint32_t Ii = 0; Ii < StepCount; ++Ii)
    {
        for(uint32_t Ai = 0; Ai < StepCount*2; ++Ai)
        {
            const float Incl = (float)Ii*InclInc, NIncl = (float)(Ii + 1)*InclInc;
            const float Azim = (float)Ai*AzimInc, NAzim = (float)(Ai + 1)*AzimInc;
            const float Blue = (1.0f + cosf(Azim))/2.0f, NBlue = (1.0f + cosf(NAzim))/2.0f;
            float CB[4], NB[4], NNB[4], CNB[4];     /* Cur+Blue, Cur+Inc+Blue, Cur+Inc+NextBlue, Cur+NextBlue */
            for(int i = 0; i < 4; ++i)
            {
                const float b = (i == 2) ? Blue : 0.0f, nb = (i == 2) ? NBlue : 0.0f;
                CB[i] = Cur[i] + b; NB[i] = (Cur[i] + Inc[i]) + b; NNB[i] = (Cur[i] + Inc[i]) + nb; CNB[i] = Cur[i] + nb;
            }
            const float P1[3] = { sinf(Incl)*cosf(Azim), cosf(Incl), sinf(Incl)*sinf(Azim) };
            const float P2[3] = { sinf(NIncl)*cosf(Azim), cosf(NIncl), sinf(NIncl)*sinf(Azim) };
            const float P3[3] = { sinf(NIncl)*cosf(NAzim), cosf(NIncl), sinf(NIncl)*sinf(NAzim) };
            const float P4[3] = { sinf(Incl)*cosf(NAzim), cosf(Incl), sinf(Incl)*sinf(NAzim) };
            if(Ii == 0)                                                     /* :4156-4189 */
            {
                sphere_vertex(0.0f, 1.0f, 0.0f, Radius, 0.5f, 0.5f, CB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(P2[0], P2[1], P2[2], Radius, P2[0], P2[2], NB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(P3[0], P3[1], P3[2], Radius, P3[0], P3[2], NNB, Pos, Col, Nrm, UV, N++);
            }
            else if(Ii == StepCount - 1)                                    /* :4190-4223 */
            {
                sphere_vertex(P1[0], P1[1], P1[2], Radius, 0.5f, 0.5f, CB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(0.0f, -1.0f, 0.0f, Radius, 0.0f, 0.0f, NB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(P4[0], P4[1], P4[2], Radius, P4[0], P4[2], NNB, Pos, Col, Nrm, UV, N++);
            }
            else                                                            /* :4224-4281 */
            {
#define ORC_UV(P) ((P)[0] + 1.0f)/2.0f, ((P)[1] + 1.0f)/2.0f
                sphere_vertex(P1[0], P1[1], P1[2], Radius, ORC_UV(P1), CB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(P2[0], P2[1], P2[2], Radius, ORC_UV(P2), NB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(P3[0], P3[1], P3[2], Radius, ORC_UV(P3), NNB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(P1[0], P1[1], P1[2], Radius, ORC_UV(P1), CB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(P3[0], P3[1], P3[2], Radius, ORC_UV(P3), NNB, Pos, Col, Nrm, UV, N++);
                sphere_vertex(P4[0], P4[1], P4[2], Radius, ORC_UV(P4), CNB, Pos, Col, Nrm, UV, N++);
#undef ORC_UV
            }
        }
        for(int i = 0; i < 4; ++i) Cur[i] = Cur[i] + Inc[i];               /* :4286 */
    }
    return N;
}
