/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's depth pre-filter, the step that produces
 * textures[DEPTH_FILTERED], the input of RGBDOdometryef::initICP (apps/elastic_fusion_file.cpp:342-351, 368).
 *
 * PARITY UNPINNED: in the reference this is a GLSL fragment shader (gl/shaders/depth_bilateral.frag:30-76) run through
 * gl/ComputePack.cpp:41-73.  There is no GL context in this environment, the reference ships no golden image of it, and
 * GLSL leaves the accuracy of exp() and of the division to the implementation, so no bit-level ground truth exists.
 * This file follows the shader statement by statement in IEEE fp32 (expf from libm); tests compare the CUDA kernel
 * with it (equal except where sum1/sum2 falls within rounding noise of a .5 boundary: +-1 mm) and with an
 * independent fp64 evaluation.
 */
#include <math.h>
#include <stdint.h>

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* depth_bilateral.frag:30-76.  in/out: dense u16 [rows][cols] millimetres; maxD in metres. */
void oracle_depth_bilateral(const uint16_t * in, int rows, int cols, float maxD, uint16_t * out)
{
    const float sigma_space2_inv_half = 0.024691358f; /* :44 */
    const float sigma_color2_inv_half = 0.000555556f; /* :45 */
    const int R = 6, D = R * 2 + 1;                   /* :47-48 */
    const unsigned cut = (unsigned)(maxD * 1000.0f);  /* :34 uint(maxD * 1000.0f) */
#pragma omp parallel for schedule(static)
    for(int y = 0; y < rows; y++)
        for(int x = 0; x < cols; x++)
        {
            const unsigned value = in[y * cols + x];
            if(value > cut || value < 300u) /* :34-37 */
            {
                out[y * cols + x] = 0;
                continue;
            }
            const int tx = imin(x - D / 2 + D, cols); /* :50 */
            const int ty = imin(y - D / 2 + D, rows); /* :51 */
            float sum1 = 0.f, sum2 = 0.f;
            for(int cy = imax(y - D / 2, 0); cy < ty; ++cy)     /* :56 */
                for(int cx = imax(x - D / 2, 0); cx < tx; ++cx) /* :58 */
                {
                    const unsigned tmp = in[cy * cols + cx];
                    const float dx = (float)x - (float)cx, dy = (float)y - (float)cy;
                    const float space2 = dx * dx + dy * dy;                            /* :65 */
                    const float dc = (float)value - (float)tmp;
                    const float color2 = dc * dc;                                      /* :66 */
                    const float weight = expf(-(space2 * sigma_space2_inv_half + color2 * sigma_color2_inv_half)); /* :68 */
                    sum1 += (float)tmp * weight; /* :70 */
                    sum2 += weight;              /* :71 */
                }
            out[y * cols + x] = (uint16_t)(unsigned)roundf(sum1 / sum2); /* :75 uint(round(sum1/sum2)) */
        }
}
