// GpuIngestionStore.cs — IIngestionStore (Services/IIngestionStore.cs:5-17) whose chunk rows live
// in HBM.  Behaviour follows InMemoryIngestionStore.cs line by line; the only additions are the
// two native calls in UpsertChunksAsync / DeleteDocumentAsync (orr_store_upsert_document_texts takes the
// Content strings as they are: tokenising, hashing and the vocabulary are the library's business).  Source only (no .NET SDK here);
// omni_recall_rag_b200/store.py is the tested mirror of this class.
using System.Collections.Concurrent;
using System.Text;
using OmniRecall.Api.Data.Models;

namespace OmniRecall.Api.Services.Gpu;

public sealed class GpuIngestionStore : IIngestionStore, IDisposable
{
    private readonly ConcurrentDictionary<string, CosmosDocumentRecord> _documents = new();
    private readonly ConcurrentDictionary<string, List<CosmosChunkRecord>> _chunksByDocument = new();
    private readonly ConcurrentDictionary<ulong, CosmosChunkRecord> _chunkByRow = new();
    private readonly ConcurrentDictionary<string, ulong[]> _rowsByDocument = new();
    private readonly object _mutate = new();
    internal nint Handle { get; }
    internal int Dim { get; }
    internal int TermSlots { get; }

    public GpuIngestionStore(IConfiguration configuration)
    {
        var cfg = new OrrConfig();
        OrrNative.orr_config_default(ref cfg);
        cfg.Device = configuration.GetValue("Gpu:Device", 0);
        cfg.Dim = Dim = configuration.GetValue("Gpu:Dim", 3072);
        cfg.TermSlots = TermSlots = configuration.GetValue("Gpu:TermSlots", 128);
        cfg.CapacityRows = configuration.GetValue("Gpu:CapacityRows", 1L << 20);
        OrrNative.Check(OrrNative.orr_store_create(in cfg, out var h));
        Handle = h;
        // keep the lower-cased Content in HBM: terms inside many vocabulary words ("ai", one letter) and chunks with more
        // distinct tokens than Gpu:TermSlots are then matched in text mode instead of being refused
        if (configuration.GetValue("Gpu:KeepText", true))
        {
            OrrNative.Check(OrrNative.orr_store_set_option(h, "text_bytes_per_row", configuration.GetValue("Gpu:TextBytesPerRow", 2048.0)));
            OrrNative.Check(OrrNative.orr_store_set_option(h, "keep_text", 1));
        }
    }

    public Task<CosmosDocumentRecord> UpsertDocumentAsync(CosmosDocumentRecord document, CancellationToken ct = default)
    {
        _documents[document.Id] = document;
        return Task.FromResult(document);
    }

    public unsafe Task UpsertChunksAsync(IReadOnlyList<CosmosChunkRecord> chunks, CancellationToken ct = default)
    {
        if (chunks.Count == 0) return Task.CompletedTask;
        var documentId = chunks[0].DocumentId;                       // InMemoryIngestionStore.cs:22
        var ordered = chunks.OrderBy(c => c.ChunkIndex).ToList();    // :23
        var n = ordered.Count;
        var emb = new float[(long)n * Dim];
        var has = new byte[n];
        var ticks = new long[n];
        // Content goes down as it is: the library lower-cases (RecallSearchService.cs:110), splits into white-space
        // tokens, hashes the distinct ones into the chunk's term set, registers them in the live vocabulary and (option
        // keep_text) keeps the lower-cased text in HBM for text mode
        var contentOffsets = new ulong[n + 1];
        using var text = new MemoryStream();
        for (var i = 0; i < n; i++)
        {
            var c = ordered[i];
            if (c.Embedding is { Count: > 0 } e && e.Count == Dim)   // other widths score cosine 0 (:71-72)
            {
                for (var j = 0; j < Dim; j++) emb[(long)i * Dim + j] = e[j];
                has[i] = 1;
            }
            ticks[i] = c.CreatedAtUtc.Ticks;
            text.Write(Encoding.UTF8.GetBytes(c.Content ?? string.Empty));
            contentOffsets[i + 1] = (ulong)text.Length;
        }
        var rows = new ulong[n];
        var textBytes = text.Length > 0 ? text.ToArray() : new byte[1];
        lock (_mutate)
        {
            fixed (float* pe = emb) fixed (byte* ph = has) fixed (long* pt = ticks) fixed (ulong* pr = rows)
            fixed (byte* ptx = textBytes) fixed (ulong* pto = contentOffsets)
                OrrNative.Check(OrrNative.orr_store_upsert_document_texts(
                    Handle, HashTerm("doc:" + documentId), n, pe, ph, pt, ptx, pto, pr));
            ForgetRows(documentId);
            _chunksByDocument[documentId] = ordered;
            _rowsByDocument[documentId] = rows;
            for (var i = 0; i < n; i++) _chunkByRow[rows[i]] = ordered[i];
        }
        return Task.CompletedTask;
    }

    public Task DeleteDocumentAsync(string documentId, CancellationToken ct = default)
    {
        _documents.TryRemove(documentId, out _);
        lock (_mutate)
        {
            if (_chunksByDocument.TryRemove(documentId, out _))
            {
                ForgetRows(documentId);
                OrrNative.Check(OrrNative.orr_store_delete_document(Handle, HashTerm("doc:" + documentId)));
            }
        }
        return Task.CompletedTask;
    }

    // The remaining six methods are InMemoryIngestionStore.cs:27-48,57-76 verbatim over the host
    // dictionaries (GetDocumentAsync, ListDocumentsAsync, GetChunksByDocumentIdAsync,
    // GetRecentChunksAsync, GetDocumentsByIdsAsync); HealthProbeService only needs
    // ListDocumentsAsync(1).
    public Task<CosmosDocumentRecord?> GetDocumentAsync(string id, CancellationToken ct = default)
        => Task.FromResult(_documents.TryGetValue(id, out var d) ? d : null);
    public Task<IReadOnlyList<CosmosDocumentRecord>> ListDocumentsAsync(int maxCount, CancellationToken ct = default)
        => Task.FromResult<IReadOnlyList<CosmosDocumentRecord>>(
            _documents.Values.OrderByDescending(d => d.CreatedAtUtc).Take(Math.Max(1, maxCount)).ToList());
    public Task<IReadOnlyList<CosmosChunkRecord>> GetChunksByDocumentIdAsync(string id, CancellationToken ct = default)
        => Task.FromResult<IReadOnlyList<CosmosChunkRecord>>(_chunksByDocument.TryGetValue(id, out var c) ? c : []);
    public Task<IReadOnlyList<CosmosChunkRecord>> GetRecentChunksAsync(int maxCount, CancellationToken ct = default)
        => Task.FromResult<IReadOnlyList<CosmosChunkRecord>>(
            _chunksByDocument.Values.SelectMany(v => v).OrderByDescending(c => c.CreatedAtUtc).Take(Math.Max(1, maxCount)).ToList());
    public Task<IReadOnlyDictionary<string, CosmosDocumentRecord>> GetDocumentsByIdsAsync(
        IReadOnlyCollection<string> ids, CancellationToken ct = default)
    {
        var set = new HashSet<string>(ids);
        return Task.FromResult<IReadOnlyDictionary<string, CosmosDocumentRecord>>(
            _documents.Where(kv => set.Contains(kv.Key)).ToDictionary(kv => kv.Key, kv => kv.Value));
    }

    internal CosmosChunkRecord ChunkOfRow(ulong row) => _chunkByRow[row];

    /// Maintenance the HBM layout needs (not part of IIngestionStore): squeeze out the rows that replace /
    /// delete tombstoned and remap the host's row tables.  Run from a background service when
    /// rows_used - count grows (every scan still reads tombstoned rows).
    public unsafe long Compact()
    {
        lock (_mutate)
        {
            var before = OrrNative.orr_store_rows_used(Handle);
            var old = new ulong[Math.Max(1, before)];
            long nLive;
            fixed (ulong* po = old) OrrNative.Check(OrrNative.orr_store_compact(Handle, po, old.Length, out nLive));
            var newOfOld = new Dictionary<ulong, ulong>();
            for (long i = 0; i < nLive; i++) newOfOld[old[i]] = (ulong)i;
            var remapped = _chunkByRow.ToDictionary(kv => newOfOld[kv.Key], kv => kv.Value);
            _chunkByRow.Clear();
            foreach (var kv in remapped) _chunkByRow[kv.Key] = kv.Value;
            foreach (var d in _rowsByDocument.Keys.ToList())
                _rowsByDocument[d] = _rowsByDocument[d].Select(r => newOfOld[r]).ToArray();
            return before - nLive;
        }
    }

    /// Snapshot / warm load of the HBM image; the host records (documents, chunks, row tables) are
    /// serialised beside it by the caller (e.g. System.Text.Json) — see omni_recall_rag_b200/store.py.
    public void SaveShard(string path) { lock (_mutate) OrrNative.Check(OrrNative.orr_store_save(Handle, path)); }
    public void LoadShard(string path) { lock (_mutate) OrrNative.Check(OrrNative.orr_store_load(Handle, path)); }

    /// distinct tokens held by live chunks; the vocabulary lives in HBM and query terms are expanded over it on the GPU
    public long VocabularySize => OrrNative.orr_store_vocab_size(Handle);

    internal static unsafe ulong HashTerm(string lower)
    {
        var bytes = Encoding.UTF8.GetBytes(lower);
        fixed (byte* p = bytes) return OrrNative.orr_hash_term(p, bytes.Length);
    }

    private void ForgetRows(string documentId)
    {
        if (!_rowsByDocument.TryRemove(documentId, out var rows)) return;
        foreach (var r in rows) _chunkByRow.TryRemove(r, out _);   // the library released the document's words itself
    }

    public void Dispose() => OrrNative.orr_store_destroy(Handle);
}
