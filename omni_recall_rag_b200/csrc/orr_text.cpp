// orr_text.cpp — text -> hashed terms: the query side of KeywordScore
// (src/OmniRecall.Api/Services/RecallSearchService.cs:95-108) and the ingest-side term-set
// builder that replaces the per-chunk `content.ToLowerInvariant().Contains(term)` scan
// (:110-111) with a hashed per-chunk term set.
//
// Equivalence with the reference's substring semantics: query terms contain no white space
// (they come out of Split) and chunk Content is white-space-delimited words
// (SlidingWindowTextChunker.cs:10-29 joins words with ' '), so `Contains(t)` is "t is a
// substring of some word".  The term SET is exactly that when no token is a proper substring
// of another (true for the fixed-width synthetic vocabulary); for natural text the host
// expands each query term into the vocabulary words that contain it and passes several
// probes for the term (orr_search's probe_term), see INTEGRATION.md.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <vector>

#include "orr_internal.h"
#include "orr_lower_table.h"

namespace {

thread_local std::string g_last_error;

struct Utf8Cursor {
    const unsigned char* p;
    const unsigned char* end;
    bool done() const { return p >= end; }
    // decodes one scalar; malformed bytes come back as themselves, one at a time
    uint32_t next(int* nbytes) {
        const unsigned c = *p;
        const int extra = c < 0x80 ? 0 : (c >> 5) == 0x6 ? 1 : (c >> 4) == 0xE ? 2 : (c >> 3) == 0x1E ? 3 : -1;
        if (extra > 0 && (end - p) > extra) {
            uint32_t cp = c & (0xFFu >> (extra + 2));
            for (int i = 1; i <= extra; ++i) cp = (cp << 6) | (p[i] & 0x3Fu);
            p += extra + 1;
            *nbytes = extra + 1;
            return cp;
        }
        p += 1;
        *nbytes = 1;
        return c;
    }
};

// char.IsWhiteSpace
bool is_space(uint32_t c) {
    switch (c) {
        case 0x20: case 0x85: case 0xA0: case 0x1680: case 0x2028: case 0x2029: case 0x202F:
        case 0x205F: case 0x3000:
            return true;
        default:
            return (c >= 0x09 && c <= 0x0D) || (c >= 0x2000 && c <= 0x200A);
    }
}

// ToLowerInvariant (RecallSearchService.cs:96,110): the Unicode simple lower-case mapping of every code point (surrogate
// pairs included), U+0130 left unchanged as .NET's invariant casing does.  Table: orr_lower_table.h (generated).
uint32_t fold(uint32_t c) {
    if (c < 0x80) return (c >= 'A' && c <= 'Z') ? c | 0x20 : c;
    int lo = 0, hi = kOrrLowerRuns - 1;
    while (lo < hi) {                                    // last run whose `first` <= c
        const int mid = (lo + hi + 1) >> 1;
        if (kOrrLower[mid].first <= c) lo = mid; else hi = mid - 1;
    }
    const OrrLowerRun& r = kOrrLower[lo];
    if (c < r.first || c > r.last || ((c - r.first) % r.stride) != 0) return c;
    return (uint32_t)((int32_t)c + r.delta);
}

void put_utf8(std::string& s, uint32_t c) {
    if (c < 0x80) s.push_back((char)c);
    else if (c < 0x800) { s.push_back((char)(0xC0 | (c >> 6))); s.push_back((char)(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) {
        s.push_back((char)(0xE0 | (c >> 12))); s.push_back((char)(0x80 | ((c >> 6) & 0x3F)));
        s.push_back((char)(0x80 | (c & 0x3F)));
    } else {
        s.push_back((char)(0xF0 | (c >> 18))); s.push_back((char)(0x80 | ((c >> 12) & 0x3F)));
        s.push_back((char)(0x80 | ((c >> 6) & 0x3F))); s.push_back((char)(0x80 | (c & 0x3F)));
    }
}

// white-space split + lower-case + ordinal-distinct, first occurrence order
std::vector<std::string> distinct_lower_tokens(const char* s, int64_t n) {
    std::vector<std::string> out;
    std::string cur;
    Utf8Cursor it{(const unsigned char*)s, (const unsigned char*)s + (n > 0 ? n : 0)};
    auto flush = [&]() {
        if (cur.empty()) return;
        bool seen = false;
        for (const auto& t : out) if (t == cur) { seen = true; break; }
        if (!seen) out.push_back(cur);
        cur.clear();
    };
    while (!it.done()) {
        const unsigned char* at = it.p;
        int nb = 0;
        const uint32_t c = it.next(&nb);
        if (nb == 1 && *at >= 0x80) { cur.push_back((char)*at); continue; }   // malformed byte
        if (is_space(c)) flush(); else put_utf8(cur, fold(c));
    }
    flush();
    return out;
}

const char* const kStopWords[] = {  // RecallSearchService.cs:13-18
    "a", "an", "and", "are", "as", "at", "be", "by", "for", "from", "how", "in", "is", "it",
    "of", "on", "or", "that", "the", "to", "was", "what", "when", "where", "which", "who", "why",
    "with"};

bool is_stop_word(const std::string& t) {
    for (const char* w : kStopWords) if (t == w) return true;
    return false;
}

}  // namespace

// ---- shared with the text-level entry points of orr_api.cu --------------------------------------------------
std::vector<std::string> orr_distinct_lower_tokens(const char* s, int64_t n) { return distinct_lower_tokens(s, n); }
bool orr_is_stop_word(const std::string& t) { return is_stop_word(t); }
// ToLowerInvariant over a whole string (white space kept): what text mode stores per chunk (:110)
std::string orr_lower_invariant(const char* s, int64_t n) {
    std::string out;
    out.reserve((size_t)(n > 0 ? n : 0));
    Utf8Cursor it{(const unsigned char*)s, (const unsigned char*)s + (n > 0 ? n : 0)};
    while (!it.done()) {
        const unsigned char* at = it.p;
        int nb = 0;
        const uint32_t c = it.next(&nb);
        if (nb == 1 && *at >= 0x80) { out.push_back((char)*at); continue; }   // malformed byte
        put_utf8(out, fold(c));
    }
    return out;
}

void orr_set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

// FNV-1a 64 over the bytes, then a 64-bit avalanche; 0 is reserved for "empty slot"
uint64_t orr_hash_bytes(const char* s, int64_t n) {
    uint64_t h = 0xCBF29CE484222325ULL;
    for (int64_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 0x100000001B3ULL; }
    h ^= h >> 33; h *= 0xFF51AFD7ED558CCDULL; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ULL; h ^= h >> 33;
    return h ? h : 1ULL;
}

extern "C" {

const char* orr_last_error(void) { return g_last_error.c_str(); }

uint64_t orr_hash_term(const char* utf8_lower, int32_t len) {
    return orr_hash_bytes(utf8_lower, len < 0 ? 0 : len);
}

int orr_tokenize_query(const char* utf8, int32_t len, uint64_t* out_hashes, int32_t cap, int32_t* n) {
    if (!n || (len > 0 && !utf8) || cap < 0 || (cap > 0 && !out_hashes)) {
        orr_set_error("orr_tokenize_query: bad argument");
        return ORR_E_INVALID;
    }
    std::vector<std::string> raw = distinct_lower_tokens(utf8, len);           // :95-98
    std::vector<const std::string*> terms;
    for (const auto& t : raw) if (!is_stop_word(t)) terms.push_back(&t);        // :103-105
    if (terms.empty()) for (const auto& t : raw) terms.push_back(&t);           // :107-108
    *n = (int32_t)terms.size();
    if ((int64_t)terms.size() > cap) {
        orr_set_error("orr_tokenize_query: %zu terms exceed capacity %d", terms.size(), cap);
        return ORR_E_INVALID;
    }
    for (size_t i = 0; i < terms.size(); ++i)
        out_hashes[i] = orr_hash_bytes(terms[i]->data(), (int64_t)terms[i]->size());
    return ORR_OK;
}

int orr_tokenize_content(const char* utf8, int32_t len, uint64_t* out_hashes, int32_t cap, int32_t* n) {
    if (!n || (len > 0 && !utf8) || cap < 0 || (cap > 0 && !out_hashes)) {
        orr_set_error("orr_tokenize_content: bad argument");
        return ORR_E_INVALID;
    }
    std::vector<std::string> toks = distinct_lower_tokens(utf8, len);           // :110 lower-cased words
    *n = (int32_t)toks.size();
    if ((int64_t)toks.size() > cap) {
        orr_set_error("orr_tokenize_content: %zu distinct tokens exceed capacity %d", toks.size(), cap);
        return ORR_E_INVALID;
    }
    for (size_t i = 0; i < toks.size(); ++i) out_hashes[i] = orr_hash_bytes(toks[i].data(), (int64_t)toks[i].size());
    return ORR_OK;
}

}  // extern "C"
