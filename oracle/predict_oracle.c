/*
 * predict_oracle.c -- TEST INFRASTRUCTURE (never linked or called by the product).
 *
 * CPU restatement of the reference's model prediction for the tracker:
 *   IndexMap::combinedPredict      src/model/IndexMap.cpp:243-341
 *     vertex stage                 src/model/shaders/splat.vert:50-87
 *     fragment stage               src/model/shaders/combo_splat.frag:33-61, decodeColor src/model/shaders/color.glsl:27-34
 *   FillIn::vertex/normal/image    src/gl/FillIn.cpp:68-198, src/gl/shaders/fill_vertex.frag:31-58, fill_normal.frag:33-51,
 *                                  fill_rgb.frag:28-36, geometry.glsl:43-61
 *
 * PARITY UNPINNED against the reference itself: there the path is GLSL on an OpenGL context (none here), the reference
 * holds no golden images of it, and GL leaves point rasterisation partly to the implementation.  The rules this file
 * adopts where GL is silent are R1-R6 below; everything else follows the shaders statement by statement in fp32
 * (compile with -ffp-contract=off).  The draw is restated the way GL executes it -- surfels in buffer order, each
 * fragment tested GL_LESS against a 24-bit depth buffer and, on success, overwriting the colour attachments.
 *
 *   R1  window position = projectPointImage; clip on projectPoint's NDC x, y in [-1, 1] (points are clipped by centre)
 *   R2  point size clamped to [1, max_point]; a pixel is covered when its centre lies in [c - s/2, c + s/2) on both axes
 *   R3  depth: clamp(gl_FragDepth, 0, 1) -> round(d * (2^24 - 1)); cleared to 2^24 - 1; GL_LESS (first drawn wins ties)
 *   R4  NaN depth => fragment dropped
 *   R5  normalize(v) = v / sqrt(dot(v, v)), mat * vec summed left to right, image bytes = bytes of int(colour),
 *       time texel = low 16 bits of uint(colTime.z)
 *   R6  fill passes fetch NEAREST with CLAMP_TO_EDGE
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct
{
    int width, height;
    float cx, cy, fx, fy;
    float max_point;
} predict_cam;

static float dot3f(const float * a, const float * b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

/* Eigen::Matrix4f::inverse() stand-in: adjugate / determinant in fp32 (Laplace expansion along 2x2 minors of row pairs). */
void predict_oracle_inverse4(const float * m, float * out)
{
    float a[4][4];
    for(int i = 0; i < 4; i++)
        for(int j = 0; j < 4; j++) a[i][j] = m[4 * i + j];
    float cof[4][4];
    for(int i = 0; i < 4; i++)
        for(int j = 0; j < 4; j++)
        {
            float s[3][3];
            int r = 0;
            for(int ii = 0; ii < 4; ii++)
            {
                if(ii == i) continue;
                int c = 0;
                for(int jj = 0; jj < 4; jj++)
                {
                    if(jj == j) continue;
                    s[r][c++] = a[ii][jj];
                }
                r++;
            }
            const float d = s[0][0] * (s[1][1] * s[2][2] - s[1][2] * s[2][1]) - s[0][1] * (s[1][0] * s[2][2] - s[1][2] * s[2][0]) +
                            s[0][2] * (s[1][0] * s[2][1] - s[1][1] * s[2][0]);
            cof[i][j] = ((i + j) & 1) ? -d : d;
        }
    const float det = a[0][0] * cof[0][0] + a[0][1] * cof[0][1] + a[0][2] * cof[0][2] + a[0][3] * cof[0][3];
    for(int i = 0; i < 4; i++)
        for(int j = 0; j < 4; j++) out[4 * i + j] = cof[j][i] / det;
}

static void project_image(const predict_cam * k, const float * p, float * u, float * v)
{
    *u = (k->fx * p[0]) / p[2] + k->cx;      /* projectPointImage, splat.vert:43-48 */
    *v = (k->fy * p[1]) / p[2] + k->cy;
}

/*
 * surfels: count x 12 floats.  Outputs (all width x height): image u8 x4, vertex f32 x4, normal f32 x4, time u16,
 * depth24 u32 (0xFFFFFF = cleared), winner i32 (-1 = none).  Returns the number of fragments generated.
 */
long long predict_oracle_combined(const float * surfels, int count, const float * tinv, const predict_cam * k, float max_depth, float conf_threshold, int time,
                                  int max_time, int time_delta, uint8_t * image, float * vertex, float * normal, uint16_t * time_tex, uint32_t * depth24,
                                  int32_t * winner)
{
    const int W = k->width, H = k->height;
    const size_t n = (size_t)W * H;
    memset(image, 0, n * 4);                 /* glClearColor(0, 0, 0, 0); glClear(COLOR | DEPTH)   IndexMap.cpp:278-280 */
    memset(vertex, 0, n * 16);
    memset(normal, 0, n * 16);
    memset(time_tex, 0, n * 2);
    for(size_t i = 0; i < n; i++)
    {
        depth24[i] = 0xFFFFFFu;
        winner[i] = -1;
    }
    const float cols = (float)W, rows = (float)H;
    long long fragments = 0;
    for(int i = 0; i < count; i++)
    {
        const float * pos = surfels + 12 * (size_t)i, * col = pos + 4, * nr = pos + 8;
        /* vec4 vPosHome = t_inv * vec4(vPosition.xyz, 1.0)   splat.vert:52 */
        float home[3];
        for(int r = 0; r < 3; r++) home[r] = ((tinv[4 * r] * pos[0] + tinv[4 * r + 1] * pos[1]) + tinv[4 * r + 2] * pos[2]) + tinv[4 * r + 3];
        /* splat.vert:54 */
        if(home[2] > max_depth || home[2] < 0.f || pos[3] < conf_threshold || (float)time - col[3] > (float)time_delta || col[3] > (float)max_time) continue;
        /* normRad = vec4(normalize(mat3(t_inv) * vNormRad.xyz), vNormRad.w)   splat.vert:65 */
        float nrm[3], rn[3];
        for(int r = 0; r < 3; r++) rn[r] = (tinv[4 * r] * nr[0] + tinv[4 * r + 1] * nr[1]) + tinv[4 * r + 2] * nr[2];
        const float len = sqrtf(dot3f(rn, rn));
        for(int r = 0; r < 3; r++) nrm[r] = rn[r] / len;
        const float rad = nr[3];
        /* gl_Position, splat.vert:61 + projectPoint :36-41 (R1) */
        float xw, yw;
        project_image(k, home, &xw, &yw);
        const float ndx = (xw - cols * 0.5f) / (cols * 0.5f), ndy = (yw - rows * 0.5f) / (rows * 0.5f);
        if(!(ndx >= -1.f && ndx <= 1.f && ndy >= -1.f && ndy <= 1.f)) continue;
        /* splat.vert:67-86 */
        float x1[3] = {nrm[1] - nrm[2], -nrm[0], nrm[0]};
        const float l1 = sqrtf(dot3f(x1, x1));
        for(int r = 0; r < 3; r++) x1[r] = ((x1[r] / l1) * rad) * 1.41421356f;
        const float y1[3] = {nrm[1] * x1[2] - nrm[2] * x1[1], nrm[2] * x1[0] - nrm[0] * x1[2], nrm[0] * x1[1] - nrm[1] * x1[0]};
        float q[4][3];
        for(int r = 0; r < 3; r++)
        {
            q[0][r] = home[r] + x1[r];
            q[1][r] = home[r] + y1[r];
            q[2][r] = home[r] - y1[r];
            q[3][r] = home[r] - x1[r];
        }
        float xmin = 0, xmax = 0, ymin = 0, ymax = 0;
        for(int c = 0; c < 4; c++)
        {
            float u, v;
            project_image(k, q[c], &u, &v);
            if(c == 0)
                xmin = xmax = u, ymin = ymax = v;
            else
                xmin = fminf(xmin, u), xmax = fmaxf(xmax, u), ymin = fminf(ymin, v), ymax = fmaxf(ymax, v);
        }
        float size = fmaxf(0.f, fmaxf(fabsf(xmax - xmin), fabsf(ymax - ymin)));
        size = fminf(fmaxf(size, 1.f), k->max_point);                              /* R2 */
        const float h = size * 0.5f;
        int x0 = (int)ceilf((xw - h) - 0.5f), x1i = (int)ceilf((xw + h) - 0.5f) - 1;
        int y0 = (int)ceilf((yw - h) - 0.5f), y1i = (int)ceilf((yw + h) - 0.5f) - 1;
        if(x0 < 0) x0 = 0;
        if(y0 < 0) y0 = 0;
        if(x1i > W - 1) x1i = W - 1;
        if(y1i > H - 1) y1i = H - 1;
        const float pn = dot3f(home, nrm), sqr_rad = rad * rad;
        for(int py = y0; py <= y1i; py++)
            for(int px = x0; px <= x1i; px++)
            {
                fragments++;
                /* combo_splat.frag:35-46 */
                const float fcx = (float)px + 0.5f, fcy = (float)py + 0.5f;
                float l[3] = {(fcx - k->cx) / k->fx, (fcy - k->cy) / k->fy, 1.f};
                const float ll = sqrtf(dot3f(l, l));
                l[0] = l[0] / ll, l[1] = l[1] / ll, l[2] = 1.f / ll;
                const float t = pn / dot3f(l, nrm);
                const float cp[3] = {t * l[0], t * l[1], t * l[2]};
                const float diff[3] = {cp[0] - home[0], cp[1] - home[1], cp[2] - home[2]};
                if(dot3f(diff, diff) > sqr_rad) continue;
                /* gl_FragDepth = (corrected_pos.z / (2 * maxDepth)) + 0.5f   combo_splat.frag:60, R3, R4 */
                float d = cp[2] / (2.f * max_depth) + 0.5f;
                if(d != d) continue;
                d = fminf(fmaxf(d, 0.f), 1.f);
                const uint32_t d24 = (uint32_t)rintf(d * 16777215.f);
                const size_t p = (size_t)py * W + px;
                if(!(d24 < depth24[p])) continue;                                /* GL_LESS */
                depth24[p] = d24;
                winner[p] = i;
                const int rgb = (int)col[0];                                     /* decodeColor, R5 */
                image[4 * p + 0] = (uint8_t)((rgb >> 16) & 0xFF);
                image[4 * p + 1] = (uint8_t)((rgb >> 8) & 0xFF);
                image[4 * p + 2] = (uint8_t)(rgb & 0xFF);
                image[4 * p + 3] = 255;
                const float z = cp[2];
                vertex[4 * p + 0] = ((fcx - k->cx) * z) * (1.f / k->fx);        /* combo_splat.frag:52 */
                vertex[4 * p + 1] = ((fcy - k->cy) * z) * (1.f / k->fy);
                vertex[4 * p + 2] = z;
                vertex[4 * p + 3] = pos[3];
                normal[4 * p + 0] = nrm[0], normal[4 * p + 1] = nrm[1], normal[4 * p + 2] = nrm[2], normal[4 * p + 3] = rad;
                time_tex[p] = (uint16_t)(uint32_t)col[2];
            }
    }
    return fragments;
}

/* getVertex(texcoord, x, y, cam, usampler2D)   geometry.glsl:43-49; cam = (cx, cy, 1/fx, 1/fy) */
static void raw_vertex(const uint16_t * depth, const predict_cam * k, int x, int y, float * out)
{
    int tx = x, ty = y;                                                          /* R6 */
    if(tx > k->width - 1) tx = k->width - 1;
    if(ty > k->height - 1) ty = k->height - 1;
    const float ifx = 1.0f / k->fx, ify = 1.0f / k->fy;
    const float z = (float)depth[(size_t)ty * k->width + tx] / 1000.0f;
    out[0] = (((float)x - k->cx) * z) * ifx;
    out[1] = (((float)y - k->cy) * z) * ify;
    out[2] = z;
}

/* which bits: 1 vertex (fill_vertex.frag), 2 normal (fill_normal.frag), 4 image (fill_rgb.frag).  Inputs for unselected passes may be NULL. */
void predict_oracle_fill(const predict_cam * k, const float * ex_vertex, const float * ex_normal, const uint8_t * ex_image, const uint16_t * raw_depth,
                         const uint8_t * raw_rgba, int which, int passthrough, float * out_vertex, float * out_normal, uint8_t * out_image)
{
    const int W = k->width, H = k->height;
    for(int y = 0; y < H; y++)
        for(int x = 0; x < W; x++)
        {
            const size_t p = (size_t)y * W + x;
            if(which & 1)
            {
                if(ex_vertex[4 * p + 2] == 0.f || passthrough)
                {
                    raw_vertex(raw_depth, k, x, y, out_vertex + 4 * p);
                    out_vertex[4 * p + 3] = 1.f;
                }
                else
                    memcpy(out_vertex + 4 * p, ex_vertex + 4 * p, 16);
            }
            if(which & 2)
            {
                if(ex_normal[4 * p + 2] == 0.f || passthrough)
                {
                    float v[3], vx[3], vy[3];
                    raw_vertex(raw_depth, k, x, y, v);
                    raw_vertex(raw_depth, k, x + 1, y, vx);
                    raw_vertex(raw_depth, k, x, y + 1, vy);
                    const float a[3] = {vx[0] - v[0], vx[1] - v[1], vx[2] - v[2]}, b[3] = {vy[0] - v[0], vy[1] - v[1], vy[2] - v[2]};
                    const float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
                    const float len = sqrtf(dot3f(c, c));
                    out_normal[4 * p + 0] = c[0] / len, out_normal[4 * p + 1] = c[1] / len, out_normal[4 * p + 2] = c[2] / len, out_normal[4 * p + 3] = 1.f;
                }
                else
                    memcpy(out_normal + 4 * p, ex_normal + 4 * p, 16);
            }
            if(which & 4)
            {
                const int sum = (int)ex_image[4 * p] + (int)ex_image[4 * p + 1] + (int)ex_image[4 * p + 2];
                memcpy(out_image + 4 * p, (sum == 0 || passthrough) ? raw_rgba + 4 * p : ex_image + 4 * p, 4);
            }
        }
}
