// fused_plan.cpp -- tiling of the dst ROI into column strips x row bands and
// the table layouts resize_fused_kernel reads.  See kernels.cu for the
// kernel and DESIGN.md for the reasoning.
#include "fused_plan.h"

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>

namespace b2r {

static int
round_taps(int n)
{
    for (int t : { 4, 6, 8, 10, 14, 16 })
        if (n <= t)
            return t;
    // wider windows (strong minification) run the variant whose tap count is
    // a run-time multiple of four
    if (n <= FUSED_MAX_HT)
        return (n + 3) & ~3;
    return 0;
}

static int
round_ring(int k)
{
    // (a ring of 2 runs as a ring of 4 with two idle slots)
    for (int t : { 4, 6, 8 })
        if (k <= t)
            return t;
    return 0;
}

// H windows of the ring-mode (downsizing) kernels: the stream kernel is
// instantiated for 8 / 10 / 14 / 16 taps; wider windows (strong minification)
// run the fused kernel's variant whose tap count is a run-time multiple of four
static int
round_taps_ring(int n)
{
    for (int t : { 8, 10, 14, 16 })
        if (n <= t)
            return t;
    if (n <= FUSED_MAX_HT)
        return (n + 3) & ~3;
    return 0;
}


// Gather-mode V tables: per dst row the first source row and vt weights.
static bool
build_v_gather(const AxisGather& gy, const std::vector<int>& first, int vt,
               int src_h, FusedPlan& P)
{
    const int H = gy.n;
    if (vt > src_h)
        return false;
    P.vk = 0;
    P.vt = vt;
    P.vfirst.resize(H);
    P.vw.assign(size_t(H) * vt, 0.0f);
    for (int k = 0; k < H; ++k) {
        int f = first[k];
        // keep first+vt inside the image: shift the window up and pad in
        // front when the run sits at the bottom edge
        int shift = 0;
        if (f + vt > src_h)
            shift = std::min(f, f + vt - src_h);
        P.vfirst[k] = f - shift;
        for (int j = 0; j < gy.count[k]; ++j)
            P.vw[size_t(k) * vt + shift + j] = gy.w[gy.woff[k] + j];
    }
    return true;
}

// Tables for resize_expand_kernel (upsizing in y).  The H pass computes a
// GROUP of four adjacent dst columns per thread from one shared window of HT
// source pixels: when upsizing in x too the four footprints nearly coincide
// (4x mitchell: 4-5 pixels for all four columns).  Strips are up to
// EXPAND_GROUPS groups wide as long as their source span fits a V row.
static bool
plan_expand(const AxisGather& gx, const AxisGather& gy,
            const std::vector<int32_t>& cfirst, const std::vector<int>& first,
            int nch, int src_h, int cta_slots, FusedPlan* plan)
{
    FusedPlan& P = *plan;
    const int W = gx.n, H = gy.n;
    if (gy.maxcount < 1 || gy.maxcount > EXPAND_MAXVT)
        return false;
    const int vt = (gy.maxcount + 1) & ~1;  // register window: 2, 4, 6 or 8 rows
    const int ngroups = (W + 3) / 4;
    int window = 1;
    for (int g = 0; g < ngroups; ++g) {
        int a = 4 * g;
        for (int k = a; k < std::min(W, a + 4); ++k)
            window = std::max(window, cfirst[k] - cfirst[a] + std::max(1, gx.count[k]));
    }
    int ht = 0;
    for (int t : { 4, 5, 6, 8 })
        if (window <= t) {
            ht = t;
            break;
        }
    if (!ht)
        return false;

    // ---- strips: greedy to find how many are needed, then balanced -------
    auto make_strips = [&](int cap, std::vector<Strip>* out, int* max_nvec) {
        out->clear();
        *max_nvec = 0;
        for (int dx0 = 0; dx0 < W;) {
            Strip s;
            s.dx0  = dx0;
            s.e0   = (cfirst[dx0] * nch) & ~3;
            s.ndx  = 0;
            s.nvec = 0;
            const int px0 = s.e0 / nch;
            while (dx0 + s.ndx < W && s.ndx < cap) {
                int g    = (dx0 + s.ndx) / 4;
                int pend = cfirst[4 * g] + ht;  // one past the last px read
                int nvec = (pend * nch - s.e0 + 3) / 4;
                if (nvec > FUSED_THREADS || pend - px0 > FUSED_THREADS)
                    break;
                s.nvec = nvec;
                s.ndx += std::min(4, W - (dx0 + s.ndx));
            }
            if (s.ndx == 0)
                return false;
            *max_nvec = std::max(*max_nvec, s.nvec);
            out->push_back(s);
            dx0 += s.ndx;
        }
        return true;
    };
    std::vector<Strip> greedy, even;
    int nv_greedy = 0, nv_even = 0;
    if (!make_strips(4 * EXPAND_GROUPS, &greedy, &nv_greedy))
        return false;
    const int n1  = int(greedy.size());
    const int cap = std::min(4 * EXPAND_GROUPS, ((W + n1 - 1) / n1 + 3) & ~3);
    if (make_strips(cap, &even, &nv_even) && int(even.size()) <= n1) {
        P.strips   = even;
        P.max_nvec = nv_even;
    } else {
        P.strips   = greedy;
        P.max_nvec = nv_greedy;
    }

    // ---- H tables ----------------------------------------------------------
    P.expand  = true;
    P.ngroups = ngroups;
    {
        // pixels between the windows of adjacent groups, on average
        double stride = ngroups > 1 ? double(cfirst[4 * (ngroups - 1)] - cfirst[0])
                                          / double(ngroups - 1)
                                    : 1.0;
        P.vrow_linear = (int(stride + 0.5) & 1) != 0;
    }
    P.ht      = ht;
    P.hfirst.resize(ngroups);
    P.hw.assign(size_t(ngroups) * ht * 4, 0.0f);
    for (int g = 0; g < ngroups; ++g) {
        int a       = 4 * g;
        P.hfirst[g] = cfirst[a];
        for (int k = a; k < std::min(W, a + 4); ++k) {
            int s = cfirst[k] - cfirst[a];
            for (int i = 0; i < gx.count[k]; ++i)
                P.hw[(size_t(s + i) * ngroups + g) * 4 + (k - a)] = gx.w[gx.woff[k] + i];
        }
    }

    // ---- V tables and bands -------------------------------------------------
    if (!build_v_gather(gy, first, vt, src_h, P))
        return false;
    const int nstrips = int(P.strips.size());
    int nbands        = std::max(1, cta_slots / std::max(1, nstrips));
    nbands            = std::min(nbands, std::max(1, H / EXPAND_BATCH));
    // a band's V weights and row offsets are staged in shared memory
    const int max_dy = 16 * 1024 / (4 * (vt + 1));
    nbands           = std::max(nbands, (H + max_dy - 1) / max_dy);
    P.bands.clear();
    P.max_band_rows = 0;
    P.max_band_dy   = 0;
    for (int b = 0; b < nbands; ++b) {
        Band bd;
        bd.dy0 = int((long long)H * b / nbands);
        int e  = int((long long)H * (b + 1) / nbands);
        bd.ndy = e - bd.dy0;
        if (bd.ndy <= 0)
            continue;
        bd.r0    = P.vfirst[bd.dy0];
        bd.nrows = P.vfirst[e - 1] + vt - bd.r0;
        bd.woff  = 0;
        bd.soff  = 0;
        P.max_band_dy   = std::max(P.max_band_dy, bd.ndy);
        P.max_band_rows = std::max(P.max_band_rows, bd.nrows);
        P.bands.push_back(bd);
    }
    return true;
}

bool
plan_fused(const AxisGather& gx, const AxisGather& gy, int nch, int src_w,
           int src_h, int cta_slots, FusedPlan* plan, int e0_align, int w2, int stream_sms,
           int nframes)
{
    // strips start on a 16-byte boundary of the source row (e0_align elements):
    // the TMA boxes of the stream kernel need it, the other kernels need 4
    e0_align = std::max(4, e0_align) & ~3;
    if (nch < 1 || nch > 4 || gx.n <= 0 || gy.n <= 0)
        return false;
    FusedPlan& P = *plan;
    P            = FusedPlan();
    P.interior_zero = gx.interior_zero || gy.interior_zero;

    const int W = gx.n, H = gy.n;
    // An output whose footprint contributes nothing (all weights zero, or only
    // black pixels) has no source index of its own: it borrows its
    // predecessor's, and LEADING ones borrow the first real one -- never 0,
    // which may lie outside the rows a host source had staged (only the span
    // the taps can touch is uploaded) and would stretch the first strip.
    int x_seed = -1, y_seed = -1;
    for (int k = 0; k < W && x_seed < 0; ++k)
        if (gx.count[k] > 0)
            x_seed = gx.first[k];
    for (int k = 0; k < H && y_seed < 0; ++k)
        if (gy.count[k] > 0)
            y_seed = gy.first[k];
    if (x_seed < 0 || y_seed < 0)
        return false;  // nothing contributes anywhere: the exact kernel writes the zeros
    std::vector<int32_t> cfirst(W);
    for (int k = 0; k < W; ++k) {
        cfirst[k] = gx.count[k] > 0 ? gx.first[k] : (k > 0 ? cfirst[k - 1] : x_seed);
        if (k > 0 && cfirst[k] < cfirst[k - 1])
            return false;  // not monotone: leave it to the exact kernel
    }
    // ---------------- V tables --------------------------------------------
    std::vector<int> first(H), last(H);
    for (int k = 0; k < H; ++k) {
        if (gy.count[k] > 0) {
            first[k] = gy.first[k];
            last[k]  = gy.first[k] + gy.count[k] - 1;
        } else {
            first[k] = last[k] = k > 0 ? last[k - 1] : y_seed;
        }
        if (k > 0) {
            if (first[k] < first[k - 1])
                return false;
            last[k] = std::max(last[k], last[k - 1]);
        }
    }
    // open rows: how many dst rows are waiting on source rows at once
    int K = 1;
    for (int k = 0, lo = 0; k < H; ++k) {
        while (last[lo] < first[k])
            ++lo;
        K = std::max(K, k - lo + 1);
    }
    const int vk          = round_ring(K);
    const long long rows_total = (long long)last[H - 1] - first[0] + 1;
    const int vt          = gy.maxcount;
    const bool ring_ok    = vk > 0;
    const bool gather_ok  = vt >= 1 && vt <= 16;
    bool use_ring         = ring_ok;
    if (ring_ok && gather_ok) {
        // FMA-quads per strip column: ring touches every source row vk times,
        // gather touches vt rows per dst row (and re-reads them through L1)
        long long ring_cost   = rows_total * vk;
        long long gather_cost = (long long)H * vt * 2;
        use_ring              = ring_cost <= gather_cost;
    }
    if (!use_ring && !gather_ok)
        return false;

    // ---------------- expand variant ---------------------------------------
    // Upsizing in y (gather mode): wide strips, a group of four dst columns
    // per H-pass thread, source rows converted to float once per strip.
    if (!use_ring && plan_expand(gx, gy, cfirst, first, nch, src_h, cta_slots, plan))
        return true;

    // ---------------- H tables --------------------------------------------
    // The H pass computes TWO adjacent dst columns per thread from one shared
    // window of source pixels (the taps of neighbouring columns overlap: 14
    // loads instead of 24 for a 2:1 lanczos3).  Per pair: the first source
    // pixel of the window and, per window tap, the weight of column A and of
    // column B (B's taps shifted by first[B]-first[A], zero where they do not
    // reach).
    const int npairs = (W + 1) / 2;
    int window = 1;
    for (int pr = 0; pr < npairs; ++pr) {
        int a = 2 * pr, bcol = a + 1;
        int span = std::max(1, gx.count[a]);
        if (bcol < W)
            span = std::max(span, cfirst[bcol] - cfirst[a] + std::max(1, gx.count[bcol]));
        window = std::max(window, span);
    }
    P.ht = use_ring ? round_taps_ring(window) : round_taps(window);
    if (!P.ht)
        return false;
    P.hfirst.resize(npairs);
    P.hw.assign(size_t(npairs) * P.ht * 2, 0.0f);
    // per pair: bit j (column A) / 16 + j (column B) set when window tap j
    // merges clamped taps of both signs (windows of <= 16 taps only)
    P.hmix.assign(npairs, 0);
    for (int pr = 0; pr < npairs; ++pr) {
        int a = 2 * pr, bcol = a + 1;
        P.hfirst[pr] = cfirst[a];
        float* w     = &P.hw[size_t(pr) * P.ht * 2];
        for (int i = 0; i < gx.count[a]; ++i) {
            w[2 * i] = gx.w[gx.woff[a] + i];
            if (i < 16 && gx.mixed[gx.woff[a] + i])
                P.hmix[pr] |= 1 << i;
        }
        if (bcol < W) {
            int s = cfirst[bcol] - cfirst[a];
            for (int i = 0; i < gx.count[bcol]; ++i) {
                w[2 * (s + i) + 1] = gx.w[gx.woff[bcol] + i];
                if (s + i < 16 && gx.mixed[gx.woff[bcol] + i])
                    P.hmix[pr] |= 1 << (16 + s + i);
            }
        }
    }

    // ---------------- column strips ---------------------------------------
    // V rows are parked as float4 pixels (1-3 channel rows padded): a row
    // holds FUSED_THREADS pixels, and a strip stages <= FUSED_THREADS
    // four-element groups.  Strips hold an even number of columns (pairs).
    // (stream kernel, w2 = 2: strips twice as wide -- 512 V threads)
    if (!(use_ring && P.ht <= 16) || w2 < 1 || w2 > 2)
        w2 = 1;
    P.w2 = w2;
    for (int dx0 = 0; dx0 < W;) {
        Strip s;
        s.dx0  = dx0;
        s.e0   = (cfirst[dx0] * nch) / e0_align * e0_align;
        s.ndx  = 0;
        s.nvec = 0;
        const int px0 = s.e0 / nch;
        while (dx0 + s.ndx < W && s.ndx < FUSED_HCOLS * w2) {
            // take columns two at a time: the pair's window end bounds both
            int pr   = (dx0 + s.ndx) / 2;
            int pend = P.hfirst[pr] + P.ht;          // one past the last px read
            int nvec = (pend * nch - s.e0 + 3) / 4;
            // (the stream kernel runs only the V warpgroups a strip of nch
            // channels can fill)
            const int vcap = std::min(FUSED_THREADS * w2, stream_vthreads(nch, w2));
            if (nvec > vcap || pend - px0 > FUSED_THREADS * w2)
                break;
            s.nvec = nvec;
            s.ndx += std::min(2, W - (dx0 + s.ndx));
        }
        if (s.ndx == 0)
            return false;  // one pair's footprint does not fit a strip
        P.max_nvec = std::max(P.max_nvec, s.nvec);
        P.strips.push_back(s);
        dx0 += s.ndx;
    }

    // ---------------- row bands -------------------------------------------
    const int nstrips = int(P.strips.size());
    int nbands        = std::max(1, cta_slots / std::max(1, nstrips));
    const int min_rows = w2 == 2 ? 32 : 8;
    nbands             = std::min(nbands, std::max(1, H / min_rows));
    // (B2R_BANDSEARCH=0: the fixed two-items-per-SM split, for A/B measurements)
    static const bool band_search = [] {
        const char* e = getenv("B2R_BANDSEARCH");
        return !(e && e[0] == '0');
    }();
    if (band_search && stream_sms > 0 && use_ring && w2 == 2) {
        // Persistent kernel: CTA c runs items c, c + sms, ... of the
        // (frame, band, strip) list, so the schedule takes ceil(items / sms)
        // rounds of one band each.  A band costs its source rows -- its share
        // of the rows plus the filter's reach into the neighbouring bands,
        // which is read again -- plus a fixed start-up (tables, pipe fill).
        // More bands fill the last round better, fewer re-read less: pick
        // the cheapest count.
        const int halo  = std::max(0, vt - int(rows_total / std::max(1, H)));
        const int setup = 12;
        const int nbmax = std::max(1, std::min(H / min_rows, 4 * stream_sms));
        // (at least one item per CTA when the shape allows it)
        const long long per_band = (long long)nstrips * std::max(1, nframes);
        const int nbmin = int(std::min<long long>(nbmax, (stream_sms + per_band - 1) / per_band));
        long long best  = -1;
        for (int nb = nbmin; nb <= nbmax; ++nb) {
            const long long items  = (long long)nstrips * nb * std::max(1, nframes);
            const long long rounds = (items + stream_sms - 1) / stream_sms;
            const long long cost   = rounds * ((rows_total + nb - 1) / nb + halo + setup);
            if (best < 0 || cost < best) {
                best   = cost;
                nbands = nb;
            }
        }
    }
    // ring weights live in shared memory: keep a band's source rows bounded
    // ring row = vk weights padded to float4 loads
    const int vkp           = (vk + 3) & ~3;
    // (wide H windows already spend shared memory on their weights: a smaller
    // ring-weight budget there keeps two CTAs resident per SM)
    const int ring_budget   = ((P.ht > 16 && P.ht <= 32) ? 10 : 24) * 1024;
    // (the stream kernel streams its ring table through the TMA stages: no bound)
    const int max_rows_smem = (use_ring && w2 == 1) ? (ring_budget / (4 * std::max(4, vkp)))
                                                    : INT_MAX;
    for (;;) {
        bool ok = true;
        P.bands.clear();
        P.max_band_rows = 0;
        P.max_band_dy   = 0;
        int woff        = 0;
        for (int b = 0; b < nbands; ++b) {
            Band bd;
            bd.dy0 = int((long long)H * b / nbands);
            int e  = int((long long)H * (b + 1) / nbands);
            bd.ndy = e - bd.dy0;
            if (bd.ndy <= 0)
                continue;
            bd.r0    = first[bd.dy0];
            bd.nrows = last[e - 1] - bd.r0 + 1;
            bd.woff  = woff;
            bd.soff  = 0;
            P.max_band_dy = std::max(P.max_band_dy, bd.ndy);
            if (use_ring) {
                if (bd.nrows > max_rows_smem) {
                    ok = false;
                    break;
                }
                woff += bd.nrows;
                P.max_band_rows = std::max(P.max_band_rows, bd.nrows);
            }
            P.bands.push_back(bd);
        }
        if (ok)
            break;
        nbands *= 2;
        if (nbands > H)
            return false;
    }

    if (use_ring) {
        P.vk = vk;
        P.vt = 0;
        P.vlast.assign(last.begin(), last.end());
        size_t total = 0;
        for (const Band& b : P.bands)
            total += size_t(b.nrows);
        P.vring.assign(total * vkp, 0.0f);
        for (const Band& b : P.bands) {
            for (int k = b.dy0; k < b.dy0 + b.ndy; ++k) {
                int slot = (k - b.dy0) % vk;
                for (int j = 0; j < gy.count[k]; ++j) {
                    int r = gy.first[k] + j;
                    P.vring[(size_t(b.woff) + (r - b.r0)) * vkp + slot]
                        = gy.w[gy.woff[k] + j];
                }
            }
        }
        // The stream kernel's view of the same weights: per source row, slot i
        // is the weight of the i-th OLDEST open destination row (the kernel
        // rotates its accumulators when a row completes instead of indexing
        // them), and the last float of the row holds, as an integer, how many
        // destination rows complete with this source row.
        const int vsw = stream_wrow(vk);
        const int batch = stream_batch(vk, P.w2);
        size_t stotal = 0;
        for (Band& b : P.bands) {
            b.soff = int(stotal);
            stotal += size_t((b.nrows + STREAM_ROWPAD - 1) / STREAM_ROWPAD) * STREAM_ROWPAD;
        }
        P.vstream.assign(stotal * vsw, 0.0f);
        for (const Band& b : P.bands) {
            int oldest = b.dy0;  // first dst row of the band not yet complete
            for (int r = b.r0; r < b.r0 + b.nrows; ++r) {
                float* row = &P.vstream[(size_t(b.soff) + (r - b.r0)) * vsw];
                while (oldest < b.dy0 + b.ndy && last[oldest] < r)
                    ++oldest;
                int done = 0, mix = 0;
                for (int k = oldest; k < std::min(oldest + vk, b.dy0 + b.ndy); ++k) {
                    if (gy.count[k] > 0 && r >= gy.first[k] && r < gy.first[k] + gy.count[k]) {
                        row[k - oldest] = gy.w[gy.woff[k] + (r - gy.first[k])];
                        if (gy.mixed[gy.woff[k] + (r - gy.first[k])])
                            mix |= 1 << (k - oldest);
                    }
                }
                for (int k = oldest; k < b.dy0 + b.ndy && last[k] == r; ++k)
                    ++done;
                // bits 0-7: rows completing here; 8-15: slots whose weight merges
                // clamped taps of both signs; 16 / 17: the completing row opens /
                // closes a hand-over batch of the band (only read when one row
                // completes; the kernel counts for itself otherwise)
                const int n = oldest - b.dy0;  // index in the band of the first completing row
                int32_t di = done | (mix << 8);
                if (done == 1) {
                    if (n % batch == 0)
                        di |= 1 << 16;
                    if (n % batch == batch - 1 || n == b.ndy - 1)
                        di |= 1 << 17;
                }
                memcpy(&row[vsw - 1], &di, sizeof(di));
            }
            // stage descriptors (slot vk of a stage's first row; stages are
            // `rows` source rows from the band's start): bit j = exactly one
            // destination row completes with row j of the stage, 0x100 = some
            // row completes several (the kernel then decodes row by row)
            const int rows    = stream_rows(P.w2);
            const int padded  = (b.nrows + STREAM_ROWPAD - 1) / STREAM_ROWPAD * STREAM_ROWPAD;
            for (int r0s = 0; r0s < padded; r0s += rows) {
                int32_t desc = 0;
                for (int j = 0; j < rows; ++j) {
                    int32_t di;
                    memcpy(&di, &P.vstream[(size_t(b.soff) + r0s + j) * vsw + vsw - 1], sizeof(di));
                    const int ne = di & 0xff;
                    if (ne == 1)
                        desc |= 1 << j;
                    else if (ne > 1)
                        desc |= 0x100;
                }
                memcpy(&P.vstream[(size_t(b.soff) + r0s) * vsw + vk], &desc, sizeof(desc));
            }
        }
    } else if (!build_v_gather(gy, first, vt, src_h, P)) {
        return false;
    }
    (void)src_w;
    return true;
}

}  // namespace b2r
