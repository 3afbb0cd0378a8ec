// ga_sketch.cu -- CountMinSketch on the device: replaces CountMinSketch.update / estimate
// (countminsketch.py:34-44) and the dict -> sketch pour of _make_sketch
// (debruijn_graph.py:181-188, 398-405; debug_graph.py:66-85).  The hash family is the
// reference's: one MurmurHash3_x86_32 (seed 0) of the window's bytes, row i indexed by
// h % primes_1_10_7[i].  Cells are 32-bit while being accumulated; ga_sketch_narrow() emits
// the reference's unsigned-16 rows and flags cells the reference would overflow on.
#include "ga_common.cuh"

namespace {

template <class K> __device__ __forceinline__ K slot_key(const Slot<K>* s);
template <> __device__ __forceinline__ u64 slot_key<u64>(const Slot<u64>* s) { return s->key; }
template <> __device__ __forceinline__ u128 slot_key<u128>(const Slot<u128>* s) {
    return ((u128)s->hi << 64) | s->lo;
}

__device__ __forceinline__ void sketch_add(const SketchView& sk, u32 h, u32 amount) {
    for (int r = 0; r < sk.rows; ++r) atomicAdd(sk.cells + sk.row_off[r] + (h % sk.width[r]), amount);
}

template <class K>
__global__ void update_table_kernel(const Slot<K>* __restrict__ table, u64 capacity, int w, int sym_bits,
                                    const u8* __restrict__ lut_g, SketchView sk) {
    __shared__ u8 lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < capacity; i += (u64)gridDim.x * blockDim.x) {
        K key = slot_key<K>(table + i);
        if (key == ga_empty_key<K>()) continue;
        sketch_add(sk, ga_murmur_key<K>(key, w, sym_bits, lut), table[i].val);
    }
}

__global__ void update_bytes_kernel(const u8* __restrict__ bytes, const u64* __restrict__ off,
                                    const u32* __restrict__ amounts, u64 n, SketchView sk) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        sketch_add(sk, ga_murmur_bytes(bytes + off[i], (u32)(off[i + 1] - off[i])), amounts[i]);
}

__global__ void estimate_bytes_kernel(const u8* __restrict__ bytes, const u64* __restrict__ off, u64 n,
                                      SketchView sk, u32* __restrict__ out) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        out[i] = ga_sketch_estimate(sk, ga_murmur_bytes(bytes + off[i], (u32)(off[i + 1] - off[i])));
}

__global__ void narrow_kernel(const u32* __restrict__ cells, u64 total, u16* __restrict__ out, u32* status) {
    bool over = false;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        u32 c = cells[i];
        over |= c > 65535u;
        out[i] = (u16)c;
    }
    if (over) atomicOr(status, GA_ST_U16_OVERFLOW);
}

int check_sketch(const char* fn, const ga_sketch* s) {
    if (!s || !s->cells || s->rows < 1 || s->rows > GA_MAX_SKETCH_ROWS) {
        ga_set_error("%s: bad sketch", fn);
        return GA_ERR_BAD_ARG;
    }
    for (int i = 0; i < s->rows; ++i)
        if (s->width[i] == 0) {
            ga_set_error("%s: zero-width sketch row", fn);
            return GA_ERR_BAD_ARG;
        }
    return GA_OK;
}

}  // namespace

extern "C" int ga_sketch_update_table(const void* table_dev, uint64_t capacity, int key_words, int k,
                                      int sym_bits, const uint8_t* lut_dev, const ga_sketch* sketch,
                                      ga_stream stream) {
    int rc = check_sketch("ga_sketch_update_table", sketch);
    if (rc) return rc;
    if (!table_dev || !lut_dev || capacity == 0 || (key_words != 1 && key_words != 2)) {
        ga_set_error("ga_sketch_update_table: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    SketchView sk = ga_sketch_view(sketch);
    unsigned grid = ga_grid(capacity, 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    cudaStream_t st = (cudaStream_t)stream;
    if (key_words == 1)
        update_table_kernel<u64><<<grid, 256, 0, st>>>((const Slot<u64>*)table_dev, capacity, k - 1, sym_bits, lut_dev, sk);
    else
        update_table_kernel<u128><<<grid, 256, 0, st>>>((const Slot<u128>*)table_dev, capacity, k - 1, sym_bits, lut_dev, sk);
    GA_LAUNCH_CHECK("sketch_update_table");
    return GA_OK;
}

extern "C" int ga_sketch_update_bytes(const uint8_t* bytes_dev, const uint64_t* offsets_dev,
                                      const uint32_t* amounts_dev, uint64_t n, const ga_sketch* sketch,
                                      ga_stream stream) {
    int rc = check_sketch("ga_sketch_update_bytes", sketch);
    if (rc) return rc;
    if (n == 0) return GA_OK;
    update_bytes_kernel<<<ga_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(
        bytes_dev, (const u64*)offsets_dev, amounts_dev, n, ga_sketch_view(sketch));
    GA_LAUNCH_CHECK("sketch_update_bytes");
    return GA_OK;
}

extern "C" int ga_sketch_estimate_bytes(const uint8_t* bytes_dev, const uint64_t* offsets_dev, uint64_t n,
                                        const ga_sketch* sketch, uint32_t* est_out_dev, ga_stream stream) {
    int rc = check_sketch("ga_sketch_estimate_bytes", sketch);
    if (rc) return rc;
    if (n == 0) return GA_OK;
    estimate_bytes_kernel<<<ga_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(
        bytes_dev, (const u64*)offsets_dev, n, ga_sketch_view(sketch), est_out_dev);
    GA_LAUNCH_CHECK("sketch_estimate_bytes");
    return GA_OK;
}

extern "C" int ga_sketch_narrow(const ga_sketch* sketch, uint16_t* rows_out_dev, uint32_t* status_dev,
                                ga_stream stream) {
    int rc = check_sketch("ga_sketch_narrow", sketch);
    if (rc) return rc;
    u64 total = 0;
    for (int i = 0; i < sketch->rows; ++i) total += sketch->width[i];
    unsigned grid = ga_grid(total, 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    narrow_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const u32*)sketch->cells, total, rows_out_dev, status_dev);
    GA_LAUNCH_CHECK("sketch_narrow");
    return GA_OK;
}
