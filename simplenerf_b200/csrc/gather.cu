// Batch assembly on the device (next-row N3): every per-ray table of a training batch gathered in ONE launch.
//
// Reference behaviour (src/data_preprocessors/DataPreprocessor01.py): load_nerf_cached_batch :572-620 and
// load_sparse_depth_cached_batch :655-700 build each of ~15 batch tensors as `-1 * ones(...)` followed by
// `t[mask] = table[indices[mask]]` -- a fill, two boolean-mask index operations (each with a device synchronisation to
// size its result) and a scatter per tensor.  Here one table entry is (source rows, destination rows, row mask or NULL,
// words per row, fill word) and   dst[i, :] = mask == NULL || mask[i] ? src[indices[i], :] : fill.
// Pure data movement (bit exact); one thread per 4-byte word, grid-stride, all tables in one launch.
#include "common.cuh"

namespace snerf {

struct GatherTable {
    const uint32_t* src[SNERF_GATHER_MAX_TABLES];
    uint32_t* dst[SNERF_GATHER_MAX_TABLES];
    const uint8_t* mask[SNERF_GATHER_MAX_TABLES];
    int words[SNERF_GATHER_MAX_TABLES];
    uint32_t fill[SNERF_GATHER_MAX_TABLES];
    int word0[SNERF_GATHER_MAX_TABLES + 1];   // prefix of words per row over the tables
    int n_tables;
};

__global__ void __launch_bounds__(256) gather_rows_kernel(const __grid_constant__ GatherTable t, const long long* __restrict__ indices,
                                                          int n_rows) {
    const int row_words = t.word0[t.n_tables];
    const long long total = (long long)n_rows * row_words;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(g / row_words), w = (int)(g % row_words);
        int k = 0;
        while (w >= t.word0[k + 1]) ++k;
        const int c = w - t.word0[k];
        const bool on = t.mask[k] == nullptr || t.mask[k][row];
        t.dst[k][(size_t)row * t.words[k] + c] = on ? t.src[k][(size_t)indices[row] * t.words[k] + c] : t.fill[k];
    }
}

}  // namespace snerf

using namespace snerf;

extern "C" int snerf_gather_rows(const snerf_gather_table* tables, int n_tables, const int64_t* indices, int n_rows, void* stream) {
    SNERF_REQUIRE(n_tables >= 1 && n_tables <= SNERF_GATHER_MAX_TABLES, "snerf_gather_rows: %d tables (1..%d)", n_tables,
                  SNERF_GATHER_MAX_TABLES);
    SNERF_REQUIRE(n_rows >= 0, "snerf_gather_rows: bad row count %d", n_rows);
    if (n_rows == 0) return SNERF_OK;
    SNERF_REQUIRE(tables && indices, "snerf_gather_rows: null pointer");
    GatherTable t{};
    int words = 0;
    for (int k = 0; k < n_tables; ++k) {
        SNERF_REQUIRE(tables[k].src && tables[k].dst, "snerf_gather_rows: table %d has a null pointer", k);
        SNERF_REQUIRE(tables[k].row_bytes >= 4 && tables[k].row_bytes % 4 == 0 && tables[k].row_bytes <= 4096,
                      "snerf_gather_rows: table %d has rows of %d bytes (a multiple of 4 up to 4096)", k, tables[k].row_bytes);
        t.src[k] = static_cast<const uint32_t*>(tables[k].src);
        t.dst[k] = static_cast<uint32_t*>(tables[k].dst);
        t.mask[k] = tables[k].mask;
        t.words[k] = tables[k].row_bytes / 4;
        t.fill[k] = tables[k].fill_bits;
        t.word0[k] = words;
        words += t.words[k];
    }
    t.word0[n_tables] = words;
    t.n_tables = n_tables;
    const long long total = (long long)n_rows * words;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t, reinterpret_cast<const long long*>(indices), n_rows);
    SNERF_LAUNCH_OK("gather_rows_kernel");
    return SNERF_OK;
}
