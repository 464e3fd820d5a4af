// GpuRecallSearchService.cs — IRecallSearchService (Services/RecallSearchService.cs:6-9) with the
// scoring loop and ordering (:26-37) behind orr_search.  Validation (:22-23), the embedding call
// (:25), the document lookup (:39) and the citation build (:41-54) are the reference's own lines.
// Source only (no .NET SDK here); omni_recall_rag_b200/recall.py is the tested mirror.
using OmniRecall.Api.Contracts;

namespace OmniRecall.Api.Services.Gpu;

public sealed class GpuRecallSearchService(GpuIngestionStore store, IEmbeddingClient embeddingClient, IConfiguration configuration)
    : IRecallSearchService
{
    public async Task<RecallSearchResponseDto> SearchAsync(string query, int topK, CancellationToken cancellationToken = default)
    {
        if (string.IsNullOrWhiteSpace(query))
            throw new ArgumentException("Query is required.", nameof(query));

        var queryEmbedding = await embeddingClient.EmbedAsync(query, cancellationToken);
        var cap = configuration.GetValue("Gpu:CandidateCap", 300);       // 300 = reference (:26); 0 = every chunk
        var mode = configuration.GetValue("Gpu:KeywordMode", 0);         // 0 auto, 1 hashed probes only, 2 text mode only

        // :26-37 in one native call: KeywordScore's query side (:95-108: split, lower-case, distinct, stop words), the
        // substring expansion of the terms over the live vocabulary (:110-111; a GPU scan of the vocabulary in HBM), the
        // fused scan, the exact fp64 re-score and the reference ordering
        var k = Math.Max(1, topK);
        var hits = new OrrHit[k];
        var q = queryEmbedding.Vector as float[] ?? queryEmbedding.Vector.ToArray();
        var qb = System.Text.Encoding.UTF8.GetBytes(query);
        int nOut;
        unsafe
        {
            fixed (byte* pqs = qb) fixed (float* pq = q) fixed (OrrHit* pout = hits)
                OrrNative.Check(OrrNative.orr_search_query(store.Handle, pqs, qb.Length, q.Length > 0 ? pq : null, q.Length,
                    DateTime.UtcNow.Ticks, topK, cap, mode, pout, out nOut));
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
