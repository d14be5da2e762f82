/*
 * Synthetic ICL-NUIM-shaped RGB-D data: an analytic "living room" -- an
 * axis-aligned 5.0 x 2.8 x 5.0 m box room with 6 boxes and 3 spheres, procedural albedo
 * (three sinusoid gratings + a 0.25 m checker, never 0) -- ray-cast through a pinhole
 * camera along a smooth Lissajous orbit.  Produces exactly the tracker's input formats:
 *   current frame : uint16 depth in mm (0 = invalid / beyond depth_max), RGBA8 colour
 *   model         : RGBA32F vertex (x,y,z,conf) and normal (nx,ny,nz,radius) maps in the
 *                   camera frame of the model pose (z == 0 = empty), RGBA8 colour
 * Test / benchmark data only; deterministic (PCG32 object placement, no libc rand).
 * Camera frame: x right, y down, z forward; poses are camera-to-world, row-major 4x4.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
    int width, height;
    float fx, fy, cx, cy;
    uint64_t seed;
    float depth_max;   /* metres; beyond => depth 0 */
    float model_max;   /* metres; beyond => empty model texel */
    float noise_mm;    /* std-dev of additive depth noise, 0 = none */
} synth_cfg;

typedef struct { double lo[3], hi[3]; } box_t;
typedef struct { double c[3], r; } sphere_t;

typedef struct
{
    synth_cfg cfg;
    box_t boxes[6];
    sphere_t spheres[3];
    double room[3];
} scene_t;

/* ---- PCG32 ---- */
typedef struct { uint64_t state, inc; } pcg_t;
static uint32_t pcg_next(pcg_t * g)
{
    uint64_t old = g->state;
    g->state = old * 6364136223846793005ULL + (g->inc | 1);
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((-rot) & 31));
}
static void pcg_seed(pcg_t * g, uint64_t seed, uint64_t seq)
{
    g->state = 0;
    g->inc = (seq << 1u) | 1u;
    pcg_next(g);
    g->state += seed;
    pcg_next(g);
}
static double pcg_unit(pcg_t * g) { return pcg_next(g) / 4294967296.0; }
static double pcg_range(pcg_t * g, double a, double b) { return a + (b - a) * pcg_unit(g); }

void * synth_create(const synth_cfg * cfg)
{
    scene_t * s = (scene_t *)calloc(1, sizeof(scene_t));
    s->cfg = *cfg;
    s->room[0] = 5.0; s->room[1] = 2.8; s->room[2] = 5.0;
    pcg_t g;
    pcg_seed(&g, cfg->seed, 54u);
    /* furniture hugs the walls so the orbit in the middle of the room stays free */
    for(int i = 0; i < 6; i++)
    {
        double w = pcg_range(&g, 0.4, 1.1), h = pcg_range(&g, 0.4, 1.3), d = pcg_range(&g, 0.4, 1.1);
        double ang = (i + pcg_range(&g, -0.3, 0.3)) * (2.0 * M_PI / 6.0);
        double cx = 2.5 + 1.95 * cos(ang), cz = 2.5 + 1.95 * sin(ang);
        s->boxes[i].lo[0] = cx - w / 2; s->boxes[i].hi[0] = cx + w / 2;
        s->boxes[i].lo[1] = 0.0;        s->boxes[i].hi[1] = h;
        s->boxes[i].lo[2] = cz - d / 2; s->boxes[i].hi[2] = cz + d / 2;
    }
    for(int i = 0; i < 3; i++)
    {
        double r = pcg_range(&g, 0.2, 0.4);
        double ang = (i + 0.5 + pcg_range(&g, -0.2, 0.2)) * (2.0 * M_PI / 3.0);
        s->spheres[i].c[0] = 2.5 + 1.6 * cos(ang);
        s->spheres[i].c[1] = pcg_range(&g, 0.9, 1.9);
        s->spheres[i].c[2] = 2.5 + 1.6 * sin(ang);
        s->spheres[i].r = r;
    }
    return s;
}

void synth_destroy(void * p) { free(p); }

/* nearest hit along o + s*d (s > eps); returns s (or -1) and the outward surface normal */
static double cast(const scene_t * sc, const double o[3], const double d[3], double n[3])
{
    double best = 1e30;
    n[0] = n[1] = n[2] = 0;
    /* room, seen from inside: first plane crossed going outwards */
    {
        double sroom = 1e30;
        int ax = -1, sign = 0;
        for(int a = 0; a < 3; a++)
        {
            if(d[a] > 1e-12)
            {
                double s = (sc->room[a] - o[a]) / d[a];
                if(s < sroom) { sroom = s; ax = a; sign = -1; }
            }
            else if(d[a] < -1e-12)
            {
                double s = (0.0 - o[a]) / d[a];
                if(s < sroom) { sroom = s; ax = a; sign = 1; }
            }
        }
        if(ax >= 0 && sroom > 1e-6)
        {
            best = sroom;
            n[0] = n[1] = n[2] = 0;
            n[ax] = sign;
        }
    }
    for(int b = 0; b < 6; b++)
    {
        const box_t * bx = &sc->boxes[b];
        double s0 = -1e30, s1 = 1e30;
        int ax0 = -1, sg0 = 0;
        int miss = 0;
        for(int a = 0; a < 3; a++)
        {
            if(fabs(d[a]) < 1e-12)
            {
                if(o[a] < bx->lo[a] || o[a] > bx->hi[a]) { miss = 1; break; }
                continue;
            }
            double ta = (bx->lo[a] - o[a]) / d[a], tb = (bx->hi[a] - o[a]) / d[a];
            int sg = -1;
            if(ta > tb) { double t = ta; ta = tb; tb = t; sg = 1; }
            if(ta > s0) { s0 = ta; ax0 = a; sg0 = sg; }
            if(tb < s1) s1 = tb;
            if(s0 > s1) { miss = 1; break; }
        }
        if(!miss && s0 > 1e-6 && s0 < best)
        {
            best = s0;
            n[0] = n[1] = n[2] = 0;
            n[ax0] = sg0;
        }
    }
    for(int k = 0; k < 3; k++)
    {
        const sphere_t * sp = &sc->spheres[k];
        double oc[3] = {o[0] - sp->c[0], o[1] - sp->c[1], o[2] - sp->c[2]};
        double A = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        double B = 2.0 * (oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2]);
        double C = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - sp->r * sp->r;
        double disc = B * B - 4 * A * C;
        if(disc <= 0) continue;
        double s = (-B - sqrt(disc)) / (2 * A);
        if(s > 1e-6 && s < best)
        {
            best = s;
            for(int a = 0; a < 3; a++) n[a] = (o[a] + s * d[a] - sp->c[a]) / sp->r;
        }
    }
    return best < 1e29 ? best : -1.0;
}

static void albedo(const double p[3], const double n[3], uint8_t rgb[3])
{
    /* three gratings + checker, per channel phase shifts; lambert-ish fixed shading per facet */
    double g1 = sin(2.0 * M_PI * (p[0] * 1.7 + p[1] * 0.6 + p[2] * 0.3));
    double g2 = sin(2.0 * M_PI * (p[0] * 0.4 - p[1] * 1.9 + p[2] * 1.1) + 1.3);
    double g3 = sin(2.0 * M_PI * (-p[0] * 9.0 + p[1] * 7.0 + p[2] * 11.0) + 2.1);
    int cx = (int)floor(p[0] / 0.25 + 1e-9), cy = (int)floor(p[1] / 0.25 + 1e-9), cz = (int)floor(p[2] / 0.25 + 1e-9);
    double chk = ((cx + cy + cz) & 1) ? 1.0 : -1.0;
    const double L[3] = {0.3, 0.8, 0.52};
    double sh = 0.75 + 0.25 * fabs(n[0] * L[0] + n[1] * L[1] + n[2] * L[2]);
    double base[3] = {128 + 30 * g1 + 22 * g2 + 20 * g3 + 36 * chk, 128 + 22 * g2 + 30 * g3 + 20 * g1 + 36 * chk, 128 + 30 * g3 + 22 * g1 + 20 * g2 + 36 * chk};
    for(int c = 0; c < 3; c++)
    {
        double v = base[c] * sh;
        int iv = (int)lrint(v);
        if(iv < 1) iv = 1;
        if(iv > 255) iv = 255;
        rgb[c] = (uint8_t)iv;
    }
}

static void ray_dir(const synth_cfg * c, const float * T, int u, int v, double dc[3], double dw[3])
{
    dc[0] = (u - (double)c->cx) / c->fx;
    dc[1] = (v - (double)c->cy) / c->fy;
    dc[2] = 1.0;
    for(int i = 0; i < 3; i++) dw[i] = T[i * 4 + 0] * dc[0] + T[i * 4 + 1] * dc[1] + T[i * 4 + 2] * dc[2];
}

/* current-frame sensor data */
void synth_render_frame(const void * p, const float * pose16, uint16_t * depth, uint8_t * rgba, uint32_t noise_seed)
{
    const scene_t * sc = (const scene_t *)p;
    const synth_cfg * c = &sc->cfg;
    const double o[3] = {pose16[3], pose16[7], pose16[11]};
#pragma omp parallel for schedule(static)
    for(int v = 0; v < c->height; v++)
    {
        pcg_t g;
        pcg_seed(&g, noise_seed, (uint64_t)v + 1);
        for(int u = 0; u < c->width; u++)
        {
            double dc[3], dw[3], n[3];
            ray_dir(c, pose16, u, v, dc, dw);
            double s = cast(sc, o, dw, n);
            const int idx = v * c->width + u;
            uint8_t rgb[3] = {1, 1, 1};
            uint16_t dmm = 0;
            if(s > 0)
            {
                double pw[3] = {o[0] + s * dw[0], o[1] + s * dw[1], o[2] + s * dw[2]};
                albedo(pw, n, rgb);
                if(s <= c->depth_max)
                {
                    double mm = s * 1000.0;
                    if(c->noise_mm > 0)
                    {
                        double u1 = pcg_unit(&g) + 1e-12, u2 = pcg_unit(&g);
                        mm += c->noise_mm * sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
                    }
                    long r = lrint(mm);
                    if(r < 1) r = 1;
                    if(r > 65535) r = 65535;
                    dmm = (uint16_t)r;
                }
            }
            depth[idx] = dmm;
            rgba[4 * idx + 0] = rgb[0];
            rgba[4 * idx + 1] = rgb[1];
            rgba[4 * idx + 2] = rgb[2];
            rgba[4 * idx + 3] = 255;
        }
    }
}

/* model prediction as the surfel renderer would hand it over (camera frame of pose16) */
void synth_render_model(const void * p, const float * pose16, float * vtx4, float * nrm4, uint8_t * rgba)
{
    const scene_t * sc = (const scene_t *)p;
    const synth_cfg * c = &sc->cfg;
    const double o[3] = {pose16[3], pose16[7], pose16[11]};
#pragma omp parallel for schedule(static)
    for(int v = 0; v < c->height; v++)
        for(int u = 0; u < c->width; u++)
        {
            double dc[3], dw[3], n[3];
            ray_dir(c, pose16, u, v, dc, dw);
            double s = cast(sc, o, dw, n);
            const int idx = v * c->width + u;
            float * V = vtx4 + 4 * idx;
            float * N = nrm4 + 4 * idx;
            uint8_t rgb[3] = {0, 0, 0};
            if(s > 0 && s <= c->model_max)
            {
                double pw[3] = {o[0] + s * dw[0], o[1] + s * dw[1], o[2] + s * dw[2]};
                albedo(pw, n, rgb);
                /* face the camera, then rotate into the camera frame (R^T n) */
                double facing = n[0] * dw[0] + n[1] * dw[1] + n[2] * dw[2];
                double sg = facing > 0 ? -1.0 : 1.0;
                double nc[3];
                for(int i = 0; i < 3; i++) nc[i] = sg * (pose16[0 * 4 + i] * n[0] + pose16[1 * 4 + i] * n[1] + pose16[2 * 4 + i] * n[2]);
                V[0] = (float)(dc[0] * s); V[1] = (float)(dc[1] * s); V[2] = (float)s; V[3] = 100.0f;
                N[0] = (float)nc[0]; N[1] = (float)nc[1]; N[2] = (float)nc[2]; N[3] = (float)(s / fabs(c->fx) * 1.41421356);
            }
            else
            {
                V[0] = V[1] = V[2] = V[3] = 0.f;
                N[0] = N[1] = N[2] = N[3] = 0.f;
            }
            rgba[4 * idx + 0] = rgb[0];
            rgba[4 * idx + 1] = rgb[1];
            rgba[4 * idx + 2] = rgb[2];
            rgba[4 * idx + 3] = 255;
        }
}

/* Smooth Lissajous orbit, <= ~1.2 cm and <= ~0.4 deg per frame at n_frames = 1000. */
void synth_trajectory(const void * p, int n_frames, uint64_t traj_seed, float * poses16)
{
    (void)p;
    pcg_t g;
    pcg_seed(&g, traj_seed, 7u);
    const double ph1 = pcg_range(&g, 0, 2 * M_PI), ph2 = pcg_range(&g, 0, 2 * M_PI), ph3 = pcg_range(&g, 0, 2 * M_PI), ph4 = pcg_range(&g, 0, 2 * M_PI);
    const double span = n_frames > 1 ? (double)(n_frames > 1000 ? n_frames : 1000) : 1000.0;
    for(int k = 0; k < n_frames; k++)
    {
        const double t = k / span;
        double pos[3] = {2.5 + 0.8 * sin(2 * M_PI * 2 * t + ph1), 1.4 + 0.3 * sin(2 * M_PI * 3 * t + ph2), 2.5 + 0.8 * sin(2 * M_PI * 1 * t + ph3)};
        const double a = 2 * M_PI * t + ph4;
        double tgt[3] = {2.5 + 2.2 * cos(a), 1.1 + 0.25 * sin(2 * M_PI * 2 * t + ph2), 2.5 + 2.2 * sin(a)};
        double f[3] = {tgt[0] - pos[0], tgt[1] - pos[1], tgt[2] - pos[2]};
        double fl = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for(int i = 0; i < 3; i++) f[i] /= fl;
        const double up[3] = {0, 1, 0};
        double r[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
        double rl = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        for(int i = 0; i < 3; i++) r[i] /= rl;
        double d[3] = {f[1] * r[2] - f[2] * r[1], f[2] * r[0] - f[0] * r[2], f[0] * r[1] - f[1] * r[0]};
        float * T = poses16 + 16 * k;
        for(int i = 0; i < 3; i++)
        {
            T[i * 4 + 0] = (float)r[i];
            T[i * 4 + 1] = (float)d[i];
            T[i * 4 + 2] = (float)f[i];
            T[i * 4 + 3] = (float)pos[i];
        }
        T[12] = T[13] = T[14] = 0.f;
        T[15] = 1.f;
    }
}
