/* Oracle (test-only): ROI Align forward, CPU, float32, NCHW -> [K,C,PH,PW].
 *
 * Restates the semantics of torchvision.ops.roi_align, the op the reference calls
 * at tracking.py:214, tracking_win.py:260, model/utils/inferScr/infer.py:163 and
 * model/utils/trainingScr/trainingCard.py:71.  The arithmetic lives in the
 * torchvision wheel (third-party, pinned torchvision==0.20.1 in
 * /root/reference/requirements.txt:2, not under /root/reference).  Rules restated
 * (SURVEY.md section 8 row R0):
 *   start = coord*scale - (aligned ? 0.5 : 0); extent = end - start, clamped to
 *   >= 1 only when !aligned; bin = extent / P; grid = sampling_ratio > 0 ?
 *   sampling_ratio : ceil(extent / P) samples per bin and axis; sample position
 *   start + p*bin + (i + .5)*bin/grid; a sample with y < -1 or y > H or x < -1 or
 *   x > W contributes 0; otherwise the coordinate is clamped to >= 0, its cell is
 *   low = (int)coord, and when low >= dim-1 both taps collapse onto dim-1; the
 *   four taps are weighted (1-ly)(1-lx), (1-ly)lx, ly(1-lx), ly*lx; the bin value
 *   is the sum over samples divided by max(grid_h*grid_w, 1).
 * Pinned in tests against the installed torchvision CPU op and tests/golden/.
 * Compile with -ffp-contract=off so coordinate arithmetic rounds like the wheel's.
 */
#include <math.h>
#include <stdint.h>

static float tap4(const float *plane, int H, int W, float y, float x)
{
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.0f;
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int y0 = (int)y, x0 = (int)x, y1, x1;
    if (y0 >= H - 1) { y0 = y1 = H - 1; y = (float)y0; } else y1 = y0 + 1;
    if (x0 >= W - 1) { x0 = x1 = W - 1; x = (float)x0; } else x1 = x0 + 1;
    float ly = y - (float)y0, lx = x - (float)x0, hy = 1.0f - ly, hx = 1.0f - lx;
    float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
    return w1 * plane[y0 * W + x0] + w2 * plane[y0 * W + x1] + w3 * plane[y1 * W + x0] +
           w4 * plane[y1 * W + x1];
}

/* feat: [B,C,H,W]; rois: [K,5] = (batch, x1, y1, x2, y2); out: [K,C,PH,PW].
 * Returns 0, or -1 when a ROI's batch index is out of range. */
int oracle_roi_align_f32(const float *feat, int B, int C, int H, int W, const float *rois,
                         int64_t K, int PH, int PW, float scale, int sampling_ratio, int aligned,
                         float *out)
{
    const float off = aligned ? 0.5f : 0.0f;
    for (int64_t k = 0; k < K; ++k) {
        const float *r = rois + 5 * k;
        int b = (int)r[0];
        if (b < 0 || b >= B) return -1;
        float sw = r[1] * scale - off, sh = r[2] * scale - off;
        float ew = r[3] * scale - off, eh = r[4] * scale - off;
        float rw = ew - sw, rh = eh - sh;
        if (!aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
        float bh = rh / (float)PH, bw = rw / (float)PW;
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
        float count = (float)(gh * gw > 1 ? gh * gw : 1);
        for (int c = 0; c < C; ++c) {
            const float *plane = feat + ((int64_t)b * C + c) * H * W;
            float *o = out + ((k * C + c) * PH) * PW;
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    float acc = 0.0f;
                    for (int iy = 0; iy < gh; ++iy) {
                        float y = sh + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
                        for (int ix = 0; ix < gw; ++ix) {
                            float x = sw + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
                            acc += tap4(plane, H, W, y, x);
                        }
                    }
                    o[ph * PW + pw] = acc / count;
                }
        }
    }
    return 0;
}
