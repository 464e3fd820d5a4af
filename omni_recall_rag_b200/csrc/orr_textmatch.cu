// orr_textmatch.cu — exact substring keyword matching on the device.
//
// KeywordScore (src/OmniRecall.Api/Services/RecallSearchService.cs:110-111) counts the query terms t with
// content.ToLowerInvariant().Contains(t, StringComparison.Ordinal).  The hashed term table of the fused scan
// answers that through a host-side expansion of t over the live vocabulary; a term that is a substring of
// more words than the kernels take probes ("ai", "go", a single letter) needs the predicate itself.  The
// store therefore keeps each chunk's lower-cased UTF-8 content in an HBM byte arena, and this kernel evaluates
// the ordinal substring test for every (row, term) into per-term row bitmaps which the exact scorer reads
// instead of the hashed table.  UTF-8 is self-synchronising, so a byte-level match of two valid strings is a
// code-point-level match, i.e. the ordinal UTF-16 Contains of the reference.
#include "orr_internal.h"

namespace {

constexpr int TM_THREADS = 256;
constexpr int TM_WIN = 1024;                                  // text bytes staged per warp and window
constexpr int TM_BUF = TM_WIN + ORR_TEXT_MAX_TERM_BYTES + 64; // window + overlap + alignment / look-ahead slack
constexpr int TM_PRE = 3;                                     // 16-byte vectors per lane that hold a row's first window in flight
static_assert(TM_PRE * 32 * 16 >= TM_BUF - 32, "the prefetch registers must cover a whole window");

struct TmWindow { const uint4* src; int head, want, n_vec; };

__device__ __forceinline__ TmWindow tm_window(const uint8_t* text, uint64_t off, int len, int base, int overlap) {
    // [base, base + TM_WIN + overlap) of the row's text, as 16-byte aligned vectors
    TmWindow w;
    w.want = max(0, min(len - base, TM_WIN + overlap));
    const uint64_t g0 = off + (uint64_t)base;
    w.head = (int)(g0 & 15u);
    w.src = reinterpret_cast<const uint4*>(text + (g0 - w.head));
    w.n_vec = w.want > 0 ? (w.head + w.want + 15) >> 4 : 0;
    return w;
}

// NT = the query's term count when it is 1..4 (patterns live in registers, the term loop is unrolled: the kernel
// is ALU-bound, ncu: 78 % ALU pipe, LOP3 a third of the instructions), 0 = any count (patterns read from smem).
template <int NT>
__global__ void __launch_bounds__(TM_THREADS, NT >= 3 ? 3 : 4) orr_text_bits_kernel(const OrrTextView tv, int64_t rows,
                                                                   const OrrTextTerms* terms_g, const uint32_t* rows_list,
                                                                   int n_list, uint32_t* bits, int64_t row_words) {
    __shared__ OrrTextTerms tt;
    __shared__ __align__(16) uint8_t wbuf[TM_THREADS / 32][TM_BUF];
    // per term: its first min(8, len) bytes as a little-endian word + mask ({pat lo, pat hi, mask lo, mask hi}, one
    // 16-byte broadcast load), so one masked 64-bit compare per (position, term) decides almost every case;
    // longer terms verify their tail byte by byte
    __shared__ uint4 t_pm[ORR_MAX_QUERY_TERMS];
    __shared__ int t_len[ORR_MAX_QUERY_TERMS];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(terms_g);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&tt);
        for (int i = threadIdx.x; i < (int)(sizeof(OrrTextTerms) / 4); i += TM_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    if (threadIdx.x < tt.n_terms) {
        const int t = threadIdx.x, o = tt.off[t], tl = tt.off[t + 1] - o;
        uint64_t pat = 0ull;
        for (int i = 0; i < min(tl, 8); ++i) pat |= (uint64_t)tt.bytes[o + i] << (8 * i);
        const uint64_t msk = tl >= 8 ? ~0ull : ((1ull << (8 * tl)) - 1ull);
        t_pm[t] = make_uint4((uint32_t)pat, (uint32_t)(pat >> 32), (uint32_t)msk, (uint32_t)(msk >> 32));
        t_len[t] = tl;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* buf = wbuf[warp];
    const int T = NT > 0 ? NT : tt.n_terms;
    uint4 pm_r[NT > 0 ? NT : 1];
    int tl_r[NT > 0 ? NT : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t) { pm_r[t] = t_pm[t]; tl_r[t] = t_len[t]; }
    const int overlap = max(tt.max_len - 1, 0);
    const int64_t gw = ((int64_t)blockIdx.x * TM_THREADS + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * TM_THREADS) >> 5;
    const int64_t n_items = rows_list ? (int64_t)n_list : rows;
    const int64_t n_blocks = (n_items + 31) >> 5;

    // matches every start position the staged window owns.  A lane takes 4 consecutive positions per step: the 11
    // bytes they cover are 4 aligned smem words, re-aligned once (the sub-word shift is the same for the whole
    // window) and funnel-shifted into the four 8-byte words; a term costs one broadcast load per step.
    auto match_window = [&](const TmWindow& w, uint64_t found) -> uint64_t {
        const uint8_t* tx = buf + w.head;
        const int starts = min(w.want, TM_WIN);
        const uint32_t sh = (uint32_t)(w.head & 3) * 8u;
        for (int p4 = lane * 4; p4 < starts; p4 += 128) {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(buf + ((w.head + p4) & ~3));
            const uint32_t x0 = wp[0], x1 = wp[1], x2 = wp[2], x3 = wp[3];
            const uint32_t a0 = __funnelshift_r(x0, x1, sh), a1 = __funnelshift_r(x1, x2, sh), a2 = __funnelshift_r(x2, x3, sh);
            uint32_t lo[4], hi[4];
            lo[0] = a0; hi[0] = a1;
#pragma unroll
            for (int j = 1; j < 4; ++j) { lo[j] = __funnelshift_r(a0, a1, 8u * j); hi[j] = __funnelshift_r(a1, a2, 8u * j); }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const uint4 pm = NT > 0 ? pm_r[NT > 0 ? t : 0] : t_pm[t];
                uint32_t hit = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) hit |= ((((lo[j] ^ pm.x) & pm.z) | ((hi[j] ^ pm.y) & pm.w)) == 0u ? 1u : 0u) << j;
                if (hit == 0u) continue;
                const int tl = NT > 0 ? tl_r[NT > 0 ? t : 0] : t_len[t];
                if (tl == 0) continue;
                for (int j = 0; j < 4; ++j) {
                    const int p = p4 + j;
                    if (!((hit >> j) & 1u) || p >= starts || p + tl > w.want) continue;   // bytes past `want` belong to no one
                    if (tl > 8) {
                        const int o = tt.off[t];
                        int i = 8;
                        while (i < tl && tx[p + i] == tt.bytes[o + i]) ++i;
                        if (i < tl) continue;
                    }
                    found |= 1ull << t;
                }
            }
        }
        return found;
    };

    for (int64_t blk = gw; blk < n_blocks; blk += W) {
        uint32_t acc_lo = 0u, acc_hi = 0u;                    // lane t: bits of the block's rows for term t / t+32
        const int n_here = (int)min((int64_t)32, n_items - (blk << 5));
        // lane r holds row r's descriptor; the first window of row r + 1 is in flight (registers) while row r is matched
        int64_t my_row = 0; unsigned long long my_off = 0ull; int my_len = 0;
        if (lane < n_here) {
            my_row = rows_list ? (int64_t)rows_list[(blk << 5) + lane] : (blk << 5) + lane;
            my_off = tv.off[my_row];
            my_len = (int)tv.len[my_row];
        }
        uint4 pre[TM_PRE];
        auto fetch_first = [&](int r) {
            const TmWindow w = tm_window(tv.text, __shfl_sync(0xffffffffu, my_off, r), __shfl_sync(0xffffffffu, my_len, r), 0, overlap);
#pragma unroll
            for (int i = 0; i < TM_PRE; ++i) {
                const int v = lane + 32 * i;
                pre[i] = v < w.n_vec ? __ldg(w.src + v) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        fetch_first(0);
        for (int r = 0; r < n_here; ++r) {
            const int64_t row = __shfl_sync(0xffffffffu, (unsigned long long)my_row, r);
            const uint64_t off = __shfl_sync(0xffffffffu, my_off, r);
            const int len = __shfl_sync(0xffffffffu, my_len, r);
            TmWindow w = tm_window(tv.text, off, len, 0, overlap);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < TM_PRE; ++i) {
                const int v = lane + 32 * i;
                if (v < w.n_vec) reinterpret_cast<uint4*>(buf)[v] = pre[i];
            }
            __syncwarp();
            if (r + 1 < n_here) fetch_first(r + 1);
            uint64_t found = match_window(w, 0ull);           // per lane: terms seen at this lane's positions
            for (int base = TM_WIN; base < len; base += TM_WIN) {
                w = tm_window(tv.text, off, len, base, overlap);
                __syncwarp();
                for (int v = lane; v < w.n_vec; v += 32) reinterpret_cast<uint4*>(buf)[v] = __ldg(w.src + v);
                __syncwarp();
                found = match_window(w, found);
            }
            const uint32_t f_lo = __reduce_or_sync(0xffffffffu, (uint32_t)found);
            const uint32_t f_hi = __reduce_or_sync(0xffffffffu, (uint32_t)(found >> 32));
            if (rows_list) {
                if (lane < T && ((f_lo >> lane) & 1u)) atomicOr(bits + (int64_t)lane * row_words + (row >> 5), 1u << (row & 31));
                if (lane + 32 < T && ((f_hi >> lane) & 1u)) atomicOr(bits + (int64_t)(lane + 32) * row_words + (row >> 5), 1u << (row & 31));
            } else {
                acc_lo |= ((f_lo >> lane) & 1u) << r;
                acc_hi |= ((f_hi >> lane) & 1u) << r;
            }
        }
        if (!rows_list) {
            if (lane < T) bits[(int64_t)lane * row_words + blk] = acc_lo;
            if (lane + 32 < T) bits[(int64_t)(lane + 32) * row_words + blk] = acc_hi;
        }
    }
}

}  // namespace

int orr_launch_text_bits(const OrrTextView& tv, int64_t rows, const OrrTextTerms* terms_dev, int n_terms, int max_len,
                         const uint32_t* rows_list, int n_list, uint32_t* bits, int64_t row_words, cudaStream_t st) {
    if (n_terms < 1 || n_terms > ORR_MAX_QUERY_TERMS || max_len > ORR_TEXT_MAX_TERM_BYTES) {
        orr_set_error("text match: %d terms / longest %d bytes exceed the limits (%d / %d)", n_terms, max_len,
                      ORR_MAX_QUERY_TERMS, ORR_TEXT_MAX_TERM_BYTES);
        return ORR_E_UNSUPPORTED;
    }
    const int64_t n_items = rows_list ? (int64_t)n_list : rows;
    if (n_items <= 0) return ORR_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t blocks_needed = ((n_items + 31) / 32 + TM_THREADS / 32 - 1) / (TM_THREADS / 32);
    const int grid = (int)std::min<int64_t>(blocks_needed, (int64_t)sms * 8);
    switch (n_terms) {
        case 1: orr_text_bits_kernel<1><<<grid, TM_THREADS, 0, st>>>(tv, rows, terms_dev, rows_list, n_list, bits, row_words); break;
        case 2: orr_text_bits_kernel<2><<<grid, TM_THREADS, 0, st>>>(tv, rows, terms_dev, rows_list, n_list, bits, row_words); break;
        case 3: orr_text_bits_kernel<3><<<grid, TM_THREADS, 0, st>>>(tv, rows, terms_dev, rows_list, n_list, bits, row_words); break;
        case 4: orr_text_bits_kernel<4><<<grid, TM_THREADS, 0, st>>>(tv, rows, terms_dev, rows_list, n_list, bits, row_words); break;
        default: orr_text_bits_kernel<0><<<grid, TM_THREADS, 0, st>>>(tv, rows, terms_dev, rows_list, n_list, bits, row_words); break;
    }
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}
