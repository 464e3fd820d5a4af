// GpuRecallSearchService.cs — IRecallSearchService (Services/RecallSearchService.cs:6-9) with the
// scoring loop and ordering (:26-37) behind orr_search.  Validation (:22-23), the embedding call
// (:25), the document lookup (:39) and the citation build (:41-54) are the reference's own lines.
// Source only (no .NET SDK here); omni_recall_rag_b200/recall.py is the tested mirror.
using OmniRecall.Api.Contracts;

namespace OmniRecall.Api.Services.Gpu;

public sealed class GpuRecallSearchService(GpuIngestionStore store, IEmbeddingClient embeddingClient, IConfiguration configuration)
    : IRecallSearchService
{
    private static readonly HashSet<string> StopWords = new(StringComparer.Ordinal)
    {
        "a", "an", "and", "are", "as", "at", "be", "by", "for", "from", "how", "in", "is",
        "it", "of", "on", "or", "that", "the", "to", "was", "what", "when", "where", "which",
        "who", "why", "with"
    };

    public async Task<RecallSearchResponseDto> SearchAsync(string query, int topK, CancellationToken cancellationToken = default)
    {
        if (string.IsNullOrWhiteSpace(query))
            throw new ArgumentException("Query is required.", nameof(query));

        var queryEmbedding = await embeddingClient.EmbedAsync(query, cancellationToken);
        var cap = configuration.GetValue("Gpu:CandidateCap", 300);       // 300 = reference (:26); 0 = every chunk

        // KeywordScore's query side (:95-108) + substring expansion over the vocabulary (:111)
        var raw = GpuIngestionStore.DistinctLowerTokens(query);
        var terms = raw.Where(t => !StopWords.Contains(t)).ToArray();
        if (terms.Length == 0) terms = raw;
        var probeHash = new List<ulong>();
        var probeTerm = new List<int>();
        for (var i = 0; i < terms.Length; i++)
            foreach (var w in store.VocabularyWordsContaining(terms[i]))
            {
                probeHash.Add(GpuIngestionStore.HashTerm(w));
                probeTerm.Add(i);
            }

        var k = Math.Max(1, topK);
        var hits = new OrrHit[k];
        var q = queryEmbedding.Vector as float[] ?? queryEmbedding.Vector.ToArray();
        int nOut;
        unsafe
        {
            if (terms.Length <= 64 && probeHash.Count <= 128)          // ORR_MAX_QUERY_TERMS / ORR_MAX_QUERY_PROBES
            {
                var ph = probeHash.Count > 0 ? probeHash.ToArray() : new ulong[1];
                var pt = probeTerm.Count > 0 ? probeTerm.ToArray() : new int[1];
                fixed (float* pq = q) fixed (ulong* pph = ph) fixed (int* ppt = pt) fixed (OrrHit* pout = hits)
                    OrrNative.Check(OrrNative.orr_search(store.Handle, q.Length > 0 ? pq : null, q.Length, terms.Length,
                        pph, ppt, probeHash.Count, DateTime.UtcNow.Ticks, topK, cap, pout, out nOut));
            }
            else
            {
                // a term is a substring of too many vocabulary words ("ai", "go", one letter): evaluate
                // Contains on the chunk text in HBM instead (slower: every candidate row is scored in fp64)
                var offs = new uint[terms.Length + 1];
                using var blob = new MemoryStream();
                for (var i = 0; i < terms.Length; i++)
                {
                    blob.Write(System.Text.Encoding.UTF8.GetBytes(terms[i]));
                    offs[i + 1] = (uint)blob.Length;
                }
                var tb = blob.Length > 0 ? blob.ToArray() : new byte[1];
                fixed (float* pq = q) fixed (byte* ptb = tb) fixed (uint* po = offs) fixed (OrrHit* pout = hits)
                    OrrNative.Check(OrrNative.orr_search_text(store.Handle, q.Length > 0 ? pq : null, q.Length, terms.Length,
                        ptb, po, DateTime.UtcNow.Ticks, topK, cap, pout, out nOut));
            }
        }

        var scored = hits.Take(nOut).Select(h => (Chunk: store.ChunkOfRow(h.Row), h.Score)).ToList();
        var documents = await store.GetDocumentsByIdsAsync(scored.Select(s => s.Chunk.DocumentId).Distinct().ToArray(), cancellationToken);
        var citations = scored.Select(s =>
        {
            documents.TryGetValue(s.Chunk.DocumentId, out var doc);
            return new RecallCitationDto(s.Chunk.DocumentId, doc?.FileName ?? "unknown", s.Chunk.Id, s.Chunk.ChunkIndex,
                TextSnippetHelper.BuildSnippet(s.Chunk.Content, 180), Math.Round(s.Score, 4), s.Chunk.CreatedAtUtc);
        }).ToList();
        return new RecallSearchResponseDto(query, citations);
    }
}
