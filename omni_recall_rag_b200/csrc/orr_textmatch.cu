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

__global__ void __launch_bounds__(TM_THREADS) orr_text_bits_kernel(const OrrTextView tv, int64_t rows,
                                                                   const OrrTextTerms* terms_g, const uint32_t* rows_list,
                                                                   int n_list, uint32_t* bits, int64_t row_words) {
    __shared__ OrrTextTerms tt;
    __shared__ __align__(16) uint8_t wbuf[TM_THREADS / 32][TM_BUF];
    // per term: its first min(8, len) bytes as a little-endian word + mask, so one 64-bit compare per
    // (position, term) decides almost every case; longer terms verify their tail byte by byte
    __shared__ uint64_t t_pat[ORR_MAX_QUERY_TERMS], t_msk[ORR_MAX_QUERY_TERMS];
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
        t_pat[t] = pat;
        t_msk[t] = tl >= 8 ? ~0ull : ((1ull << (8 * tl)) - 1ull);
        t_len[t] = tl;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* buf = wbuf[warp];
    const int T = tt.n_terms;
    const int overlap = max(tt.max_len - 1, 0);
    const int64_t gw = ((int64_t)blockIdx.x * TM_THREADS + threadIdx.x) >> 5;
    const int64_t W = ((int64_t)gridDim.x * TM_THREADS) >> 5;
    const int64_t n_items = rows_list ? (int64_t)n_list : rows;
    const int64_t n_blocks = (n_items + 31) >> 5;
    for (int64_t blk = gw; blk < n_blocks; blk += W) {
        uint32_t acc_lo = 0u, acc_hi = 0u;                    // lane t: bits of the block's rows for term t / t+32
        const int n_here = (int)min((int64_t)32, n_items - (blk << 5));
        for (int r = 0; r < n_here; ++r) {
            const int64_t row = rows_list ? (int64_t)rows_list[(blk << 5) + r] : (blk << 5) + r;
            const uint64_t off = tv.off[row];
            const int len = (int)tv.len[row];
            uint64_t found = 0ull;                            // per lane: terms seen at this lane's positions
            for (int base = 0; base < len; base += TM_WIN) {
                // stage [base, base + TM_WIN + overlap) of the row's text, 16-byte aligned loads
                const int want = min(len - base, TM_WIN + overlap);
                const uint64_t g0 = off + (uint64_t)base;
                const int head = (int)(g0 & 15u);
                const uint4* src = reinterpret_cast<const uint4*>(tv.text + (g0 - head));
                const int n_vec = (head + want + 15) >> 4;
                __syncwarp();
                for (int v = lane; v < n_vec; v += 32) reinterpret_cast<uint4*>(buf)[v] = __ldg(src + v);
                __syncwarp();
                const uint8_t* tx = buf + head;
                const int starts = min(want, TM_WIN);         // start positions owned by this window
                for (int p = lane; p < starts; p += 32) {
                    // the 8 text bytes at p as one word: three aligned 32-bit smem loads + two funnel shifts
                    const int at = head + p;
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(buf + (at & ~3));
                    const uint32_t sh = (uint32_t)(at & 3) * 8u;
                    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
                    const uint64_t w = ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | (uint64_t)__funnelshift_r(w0, w1, sh);
                    for (int t = 0; t < T; ++t) {
                        if (((w ^ t_pat[t]) & t_msk[t]) != 0ull) continue;
                        const int tl = t_len[t];
                        if (tl == 0 || p + tl > want) continue;                // bytes past `want` belong to no one
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
    orr_text_bits_kernel<<<grid, TM_THREADS, 0, st>>>(tv, rows, terms_dev, rows_list, n_list, bits, row_words);
    ORR_CUDA_OK(cudaGetLastError());
    return ORR_OK;
}
