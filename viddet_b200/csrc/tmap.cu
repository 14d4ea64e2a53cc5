// Host-side TMA tensor-map encoding.  cuTensorMapEncodeTiled is resolved through the CUDA runtime
// (cudaGetDriverEntryPoint), so the library has no link-time dependency on libcuda.
#include "tc.cuh"

namespace vd {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 tensor, innermost dimension contiguous, 128-byte swizzle (box[0] must be 64 elements),
// zero fill for out-of-bounds elements.
int encode_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return set_error(VD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
    cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(gptr), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(VD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)",
                         (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1);
    return VD_OK;
}

}  // namespace vd
