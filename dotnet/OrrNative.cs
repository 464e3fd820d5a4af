// OrrNative.cs — P/Invoke binding of liborr.so (include/orr.h), ABI version 1.
// Source only: there is no .NET SDK in the build image, so this file is not compiled here.
using System.Runtime.InteropServices;

namespace OmniRecall.Api.Services.Gpu;

[StructLayout(LayoutKind.Sequential)]
internal struct OrrConfig
{
    public int AbiVersion, Device, Dim, TermSlots;
    public long CapacityRows;
    public ulong RowBase;
    public double WCos, WKw, WRec, RecencyDays;
}

[StructLayout(LayoutKind.Sequential)]
internal struct OrrHit
{
    public ulong Row;
    public double Score;
    public long CreatedTicks;
}

internal static partial class OrrNative
{
    private const string Lib = "orr";   // liborr.so on the library path

    [LibraryImport(Lib)] internal static partial void orr_config_default(ref OrrConfig cfg);
    [LibraryImport(Lib)] internal static partial int orr_store_create(in OrrConfig cfg, out nint store);
    [LibraryImport(Lib)] internal static partial void orr_store_destroy(nint store);
    [LibraryImport(Lib)] internal static unsafe partial int orr_store_upsert_document_chunks(
        nint store, ulong docKey, int n, float* emb, byte* hasEmb, long* createdTicks,
        ulong* termHashes, uint* termOffsets, ulong* outRows);
    [LibraryImport(Lib)] internal static unsafe partial int orr_store_upsert_document_chunks_text(
        nint store, ulong docKey, int n, float* emb, byte* hasEmb, long* createdTicks,
        ulong* termHashes, uint* termOffsets, byte* textLowerUtf8, ulong* textOffsets, ulong* outRows);
    // text-level ingest: Content strings in; the library lower-cases, tokenises, hashes and keeps the live vocabulary
    [LibraryImport(Lib)] internal static unsafe partial int orr_store_upsert_document_texts(
        nint store, ulong docKey, int n, float* emb, byte* hasEmb, long* createdTicks,
        byte* contentsUtf8, ulong* contentOffsets, ulong* outRows);
    [LibraryImport(Lib)] internal static partial long orr_store_vocab_size(nint store);
    // the query STRING in: tokenising, stop words, vocabulary expansion on the GPU and the search in one call
    [LibraryImport(Lib)] internal static unsafe partial int orr_search_query(
        nint store, byte* queryUtf8, int queryLen, float* q, int qDim, long nowTicks, int topK, int candidateCap,
        int keywordMode, OrrHit* hits, out int nOut);
    [LibraryImport(Lib)] internal static partial int orr_store_delete_document(nint store, ulong docKey);
    [LibraryImport(Lib)] internal static unsafe partial int orr_store_compact(nint store, ulong* oldRowsOut, long outCap, out long nLive);
    [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] internal static partial int orr_store_save(nint store, string path);
    [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] internal static partial int orr_store_load(nint store, string path);
    [LibraryImport(Lib, StringMarshalling = StringMarshalling.Utf8)] internal static partial int orr_store_set_option(nint store, string name, double value);
    [LibraryImport(Lib)] internal static partial long orr_store_rows_used(nint store);
    [LibraryImport(Lib)] internal static partial long orr_store_count(nint store);
    [LibraryImport(Lib)] internal static unsafe partial ulong orr_hash_term(byte* utf8Lower, int len);
    [LibraryImport(Lib)] internal static unsafe partial int orr_search(
        nint store, float* q, int qDim, int nTerms, ulong* probeHash, int* probeTerm, int nProbes,
        long nowTicks, int topK, int candidateCap, OrrHit* hits, out int nOut);
    // text mode: KeywordScore's Contains evaluated on the chunk text kept in HBM (no probe limit)
    [LibraryImport(Lib)] internal static unsafe partial int orr_search_text(
        nint store, float* q, int qDim, int nTerms, byte* termsLowerUtf8, uint* termOffsets,
        long nowTicks, int topK, int candidateCap, OrrHit* hits, out int nOut);
    [LibraryImport(Lib)] internal static unsafe partial int orr_search_batch(
        nint store, int batch, float* q, int qDim, int* nTerms, ulong* probeHash, int* probeTerm, uint* probeOffsets,
        long nowTicks, int topK, OrrHit* hits, int* nOut);
    // one host process driving several GPUs: per-GPU stores + the fused peer-memory all-gather/merge
    [LibraryImport(Lib)] internal static unsafe partial int orr_search_device(
        nint store, float* qDev, int qDim, int nTerms, ulong* probeHash, int* probeTerm, int nProbes,
        long nowTicks, int topK, OrrHit* outDev, int* statusDev, nint cudaStream);
    [LibraryImport(Lib)] internal static partial int orr_xchg_create(int device, int world, int rank, int maxTopK, out nint xchg);
    [LibraryImport(Lib)] internal static partial int orr_xchg_attach_peer(nint xchg, int peerRank, nint peer);
    [LibraryImport(Lib)] internal static unsafe partial int orr_xchg_allgather_merge(
        nint xchg, OrrHit* hitsDev, int* statusDev, int topK, OrrHit* outDev, int* outStatusDev, nint cudaStream);
    [LibraryImport(Lib)] internal static partial void orr_xchg_destroy(nint xchg);
    // the whole box from this one process: one shard per GPU, fused peer-memory all-gather + merge per query
    [LibraryImport(Lib)] internal static unsafe partial int orr_cluster_create(in OrrConfig cfg, int* devices, int nDevices, int maxTopK, out nint cluster);
    [LibraryImport(Lib)] internal static partial void orr_cluster_destroy(nint cluster);
    [LibraryImport(Lib)] internal static partial long orr_cluster_count(nint cluster);
    [LibraryImport(Lib)] internal static unsafe partial int orr_cluster_upsert_document_chunks(
        nint cluster, ulong docKey, int n, float* emb, byte* hasEmb, long* createdTicks,
        ulong* termHashes, uint* termOffsets, byte* textLowerUtf8, ulong* textOffsets, ulong* outRows);
    [LibraryImport(Lib)] internal static partial int orr_cluster_delete_document(nint cluster, ulong docKey);
    [LibraryImport(Lib)] internal static unsafe partial int orr_cluster_search(
        nint cluster, float* q, int qDim, int nTerms, ulong* probeHash, int* probeTerm, int nProbes,
        long nowTicks, int topK, OrrHit* hits, out int nOut);
    // throughput forms over the cluster: a pipelined run of single queries (three in flight) and whole batches
    [LibraryImport(Lib)] internal static unsafe partial int orr_cluster_search_many(
        nint cluster, int nQueries, float* q, int qDim, int* nTerms, ulong* probeHash, int* probeTerm, uint* probeOffsets,
        long nowTicks, int topK, OrrHit* hits, int* nOut);
    [LibraryImport(Lib)] internal static unsafe partial int orr_cluster_search_batch(
        nint cluster, int batch, float* q, int qDim, int* nTerms, ulong* probeHash, int* probeTerm, uint* probeOffsets,
        long nowTicks, int topK, OrrHit* hits, int* nOut);
    [LibraryImport(Lib)] internal static partial int orr_cluster_size(nint cluster);
    [LibraryImport(Lib)] internal static partial nint orr_cluster_shard(nint cluster, int i);   // borrowed orr_store: options, snapshot, compaction
    // bulk ingest (warm load / hydration): many documents per call, searches keep running meanwhile
    [LibraryImport(Lib)] internal static unsafe partial int orr_store_upsert_documents_texts(
        nint store, int nDocs, ulong* docKeys, uint* docChunkOffsets, float* emb, byte* hasEmb, long* createdTicks,
        byte* contentsUtf8, ulong* contentOffsets, ulong* outRows);
    [LibraryImport(Lib)] internal static unsafe partial int orr_merge_hits(
        OrrHit* lists, int* listLen, int nLists, int listStride, int topK, OrrHit* hits, out int nOut);
    [LibraryImport(Lib)] internal static partial nint orr_last_error();

    internal static void Check(int rc)
    {
        if (rc == 0) return;
        var msg = Marshal.PtrToStringUTF8(orr_last_error()) ?? "liborr error";
        // surfaces through the global exception handler as HTTP 500 (Program.cs:77-99)
        throw new InvalidOperationException($"liborr {rc}: {msg}");
    }
}
