// ga_ingest.cu -- read ingestion on the device: ASCII -> 2-bit / 8-bit packed words
// (replaces the Python string list built by IOHandler.read_input, assemble.py:40-71), and a
// seeded synthetic read generator for shapes generate_reads.py:42-74 cannot make.
#include "ga_common.cuh"

namespace {

template <int SB>
__device__ __forceinline__ u64 pack_word(const u8* __restrict__ src, u32 n, const u8* __restrict__ lut,
                                         bool& bad) {
    constexpr u32 SPW = 64 / SB;
    u64 word = 0;
    for (u32 j = 0; j < SPW; ++j) {
        if (j < n) {
            u32 code = lut[src[j]];
            if (code == 0xFFu) {
                bad = true;
                code = 0;
            }
            word |= (u64)code << (SB * j);
        }
    }
    return word;
}

// uniform reads: one thread per output word
template <int SB>
__global__ void pack_uniform_kernel(const u8* __restrict__ ascii, u64 n_reads, u32 len,
                                    const u8* __restrict__ lut_g, u64* __restrict__ out,
                                    u32 stride_words, u32* status) {
    constexpr u32 SPW = 64 / SB;
    __shared__ u8 lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    u64 total = n_reads * (u64)stride_words;
    bool bad = false;
    for (u64 idx = blockIdx.x * (u64)blockDim.x + threadIdx.x; idx < total;
         idx += (u64)gridDim.x * blockDim.x) {
        u64 r = idx / stride_words;
        u32 wi = (u32)(idx % stride_words);
        u32 first = wi * SPW;
        u32 n = first < len ? (len - first < SPW ? len - first : SPW) : 0;
        out[idx] = n ? pack_word<SB>(ascii + r * (u64)len + first, n, lut, bad) : 0ull;
    }
    if (bad) atomicOr(status, GA_ST_BAD_SYMBOL);
}

// ragged reads: one thread per read
template <int SB>
__global__ void pack_ragged_kernel(const u8* __restrict__ ascii, const u64* __restrict__ in_off,
                                   u64 n_reads, const u8* __restrict__ lut_g, u64* __restrict__ out,
                                   const u64* __restrict__ out_off, u32* status) {
    constexpr u32 SPW = 64 / SB;
    __shared__ u8 lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    bool bad = false;
    for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < n_reads;
         r += (u64)gridDim.x * blockDim.x) {
        u64 b0 = in_off[r];
        u64 len = in_off[r + 1] - b0;
        u64* dst = out + out_off[r];
        for (u64 first = 0; first < len; first += SPW) {
            u32 n = (u32)(len - first < SPW ? len - first : SPW);
            dst[first / SPW] = pack_word<SB>(ascii + b0 + first, n, lut, bad);
        }
    }
    if (bad) atomicOr(status, GA_ST_BAD_SYMBOL);
}

// packed uniform reads -> ASCII (inverse of pack_uniform_kernel): one thread per packed word
template <int SB>
__global__ void unpack_uniform_kernel(const u64* __restrict__ words, u64 n_reads, u32 len, u32 stride_words,
                                      const u8* __restrict__ inv_g, u8* __restrict__ ascii) {
    constexpr u32 SPW = 64 / SB;
    constexpr u64 SMASK = (1ull << SB) - 1;
    __shared__ u8 inv[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) inv[i] = inv_g[i];
    __syncthreads();
    u64 total = n_reads * (u64)stride_words;
    for (u64 idx = blockIdx.x * (u64)blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        u64 r = idx / stride_words;
        u32 first = (u32)(idx % stride_words) * SPW;
        u64 word = words[idx];
        u8* dst = ascii + r * (u64)len + first;
        for (u32 j = 0; j < SPW && first + j < len; ++j) {
            dst[j] = inv[word & SMASK];
            word >>= SB;
        }
    }
}

__device__ __forceinline__ u64 splitmix64(u64 x) {
    u64 z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void gen_genome_kernel(u8* __restrict__ genome, u64 size, u64 seed) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < size; i += (u64)gridDim.x * blockDim.x)
        genome[i] = (u8)(splitmix64((seed << 40) + i) & 3ull);
}

// one thread per output word (32 bases)
__global__ void gen_reads_kernel(const u8* __restrict__ genome, u64 gsize, u64 first_read, u64 n_reads,
                                 u32 len, u64 seed, u32 sub_per_10k, u64* __restrict__ out,
                                 u32 stride_words, int paired, u32 mate_distance) {
    u64 total = n_reads * (u64)stride_words;
    for (u64 idx = blockIdx.x * (u64)blockDim.x + threadIdx.x; idx < total;
         idx += (u64)gridDim.x * blockDim.x) {
        u64 local = idx / stride_words;
        u32 wi = (u32)(idx % stride_words);
        u64 r = first_read + local;
        // paired: stored read r is mate (r & 1) of pair (r >> 1); mate 2 starts mate_distance later
        u64 draw = paired ? (r >> 1) : r;
        u64 start = (splitmix64((seed << 40) + (1ull << 39) + draw) % gsize +
                     (paired && (r & 1) ? mate_distance : 0u)) % gsize;
        u64 word = 0;
        u32 first = wi * 32u;
        u64 pos = (start + first) % gsize;
        for (u32 j = 0; j < 32u && first + j < len; ++j) {
            u64 h = splitmix64(((seed + 1ull) << 40) ^ (r * (u64)len + first + j));
            u32 code = genome[pos];
            if (h % 10000ull < sub_per_10k) code = (u32)(h >> 32) & 3u;
            word |= (u64)code << (2 * j);
            if (++pos == gsize) pos = 0;
        }
        out[idx] = word;
    }
}

}  // namespace

extern "C" int ga_pack_reads(const uint8_t* ascii_dev, const uint64_t* in_offsets_dev, uint64_t n_reads,
                             uint32_t uniform_len, const uint8_t* lut_dev, int storage_bits,
                             void* words_dev, const uint64_t* out_offsets_dev, uint32_t stride_words,
                             uint32_t* status_dev, ga_stream stream) {
    if (storage_bits != 2 && storage_bits != 8) {
        ga_set_error("ga_pack_reads: storage_bits must be 2 or 8");
        return GA_ERR_BAD_ARG;
    }
    if (n_reads == 0) return GA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned block = 256;
    if (in_offsets_dev == nullptr) {
        u64 total = n_reads * (u64)stride_words;
        if (total == 0) return GA_OK;
        unsigned grid = ga_grid(total, block);
        if (storage_bits == 2)
            pack_uniform_kernel<2><<<grid, block, 0, st>>>(ascii_dev, n_reads, uniform_len, lut_dev,
                                                          (u64*)words_dev, stride_words, status_dev);
        else
            pack_uniform_kernel<8><<<grid, block, 0, st>>>(ascii_dev, n_reads, uniform_len, lut_dev,
                                                          (u64*)words_dev, stride_words, status_dev);
    } else {
        if (out_offsets_dev == nullptr) {
            ga_set_error("ga_pack_reads: ragged input needs out_offsets_dev");
            return GA_ERR_BAD_ARG;
        }
        unsigned grid = ga_grid(n_reads, block);
        if (storage_bits == 2)
            pack_ragged_kernel<2><<<grid, block, 0, st>>>(ascii_dev, (const u64*)in_offsets_dev, n_reads,
                                                         lut_dev, (u64*)words_dev,
                                                         (const u64*)out_offsets_dev, status_dev);
        else
            pack_ragged_kernel<8><<<grid, block, 0, st>>>(ascii_dev, (const u64*)in_offsets_dev, n_reads,
                                                         lut_dev, (u64*)words_dev,
                                                         (const u64*)out_offsets_dev, status_dev);
    }
    GA_LAUNCH_CHECK("pack");
    return GA_OK;
}

extern "C" int ga_unpack_reads(const void* words_dev, uint64_t n_reads, uint32_t uniform_len, uint32_t stride_words,
                               int storage_bits, const uint8_t* inv_lut_dev, uint8_t* ascii_dev, ga_stream stream) {
    if ((storage_bits != 2 && storage_bits != 8) || !inv_lut_dev) {
        ga_set_error("ga_unpack_reads: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    u64 total = n_reads * (u64)stride_words;
    if (total == 0 || uniform_len == 0) return GA_OK;
    unsigned grid = ga_grid(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (storage_bits == 2)
        unpack_uniform_kernel<2><<<grid, 256, 0, st>>>((const u64*)words_dev, n_reads, uniform_len, stride_words, inv_lut_dev, ascii_dev);
    else
        unpack_uniform_kernel<8><<<grid, 256, 0, st>>>((const u64*)words_dev, n_reads, uniform_len, stride_words, inv_lut_dev, ascii_dev);
    GA_LAUNCH_CHECK("unpack");
    return GA_OK;
}

extern "C" int ga_gen_genome(uint8_t* genome_codes_dev, uint64_t size, uint64_t seed, ga_stream stream) {
    if (size == 0) return GA_OK;
    gen_genome_kernel<<<ga_grid(size, 256), 256, 0, (cudaStream_t)stream>>>(genome_codes_dev, size, seed);
    GA_LAUNCH_CHECK("gen_genome");
    return GA_OK;
}

extern "C" int ga_gen_reads(const uint8_t* genome_codes_dev, uint64_t genome_size, uint64_t first_read,
                            uint64_t n_reads, uint32_t read_len, uint64_t seed, uint32_t sub_per_10k,
                            void* words_dev, uint32_t stride_words, int paired, uint32_t mate_distance,
                            ga_stream stream) {
    if (genome_size == 0 || stride_words * 32u < read_len) {
        ga_set_error("ga_gen_reads: empty genome or stride too small");
        return GA_ERR_BAD_ARG;
    }
    if (n_reads == 0) return GA_OK;
    u64 total = n_reads * (u64)stride_words;
    gen_reads_kernel<<<ga_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        genome_codes_dev, genome_size, first_read, n_reads, read_len, seed, sub_per_10k,
        (u64*)words_dev, stride_words, paired, mate_distance);
    GA_LAUNCH_CHECK("gen_reads");
    return GA_OK;
}
