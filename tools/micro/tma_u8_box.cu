// micro-probe: one 32x32-byte box of a u8 3-D tensor map fetched by one lane (the K2 TMA A/B's request shape)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Maps { CUtensorMap map[3]; };
__global__ void probe(const __grid_constant__ Maps T, int level, int ox, int oy, int fz, uint8_t* out, int lane_sel)
{
    __shared__ __align__(128) uint8_t tile[2][1024];
    __shared__ __align__(8) unsigned long long bar[2];
    const int lane = threadIdx.x & 31;
    const int slot = lane >> 4, hl = lane & 15;
    if (hl == 0) {
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar[slot]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (hl == 0 && (lane_sel < 0 || slot == lane_sel)) {
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar[slot]);
        unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[slot][0]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(1024u) : "memory");
        const CUtensorMap* m = level == 0 ? &T.map[0] : level == 1 ? &T.map[1] : &T.map[2];
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(m), "r"(ox + slot), "r"(oy), "r"(fz), "r"(b) : "memory");
    }
    if (lane_sel < 0 || slot == lane_sel) {
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar[slot]);
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b), "r"(0u) : "memory");
        for (int i = hl; i < 1024; i += 16) out[slot * 1024 + i] = tile[slot][i];
    }
}
__global__ void probe1(const __grid_constant__ CUtensorMap M, int ox, int oy, int fz, uint8_t* out)
{
    __shared__ __align__(128) uint8_t tile[1024];
    __shared__ __align__(8) unsigned long long bar;
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(1024u) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(&M), "r"(ox), "r"(oy), "r"(fz), "r"(b) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b), "r"(0u) : "memory");
    for (int i = threadIdx.x; i < 1024; i += 32) out[i] = tile[i];
}
int main(int argc, char** argv)
{
    const int mode = argc > 3 ? atoi(argv[3]) : 0, OX = argc > 4 ? atoi(argv[4]) : 17, BW = argc > 5 ? atoi(argv[5]) : 32;
    const int PROMO = argc > 6 ? atoi(argv[6]) : 0;
    const int w = 480, h = 270, pitch = argc > 1 ? atoi(argv[1]) : 480, nf = 4;
    const long long fs = argc > 2 ? atoll(argv[2]) : (long long)pitch * h + 0;
    uint8_t* img; cudaMalloc(&img, fs * nf + 4096);
    uint8_t* himg = (uint8_t*)malloc(fs * nf);
    for (long long i = 0; i < fs * nf; ++i) himg[i] = (uint8_t)((i * 2654435761u) >> 13);
    cudaMemcpy(img, himg, fs * nf, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncFn enc = (EncFn)p;
    Maps T;
    for (int l = 0; l < 3; ++l) {
        cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nf};
        cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)fs};
        cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)(1024 / BW), 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&T.map[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode level %d -> %d\n", l, (int)r);
    }
    uint8_t* out; cudaMalloc(&out, 2048);
    uint8_t hout[2048];
    if (mode == 1) {
        probe1<<<1, 32>>>(T.map[0], OX, 33, 2, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("single map ox %d bw %d: %s\n", OX, BW, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hout, out, 1024, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < 1024 / BW; ++y) for (int x = 0; x < BW; ++x) bad += hout[y * BW + x] != himg[2 * fs + (long long)(33 + y) * pitch + OX + x];
        printf("single map ox %d bw %d promo %d: ok, mismatches %d\n", OX, BW, PROMO, bad);
        return 0;
    }
    for (int sel = 0; sel <= 2; ++sel) {
        const int lane_sel = sel == 2 ? -1 : sel;
        for (int level = 0; level < 3; ++level) {
            probe<<<1, 32>>>(T, level, OX, 33, 2, out, lane_sel);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("lane_sel %d level %d: %s\n", lane_sel, level, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hout, out, 2048, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int s = 0; s < 2; ++s) {
                if (lane_sel >= 0 && s != lane_sel) continue;
                for (int y = 0; y < 32; ++y) for (int x = 0; x < 32; ++x)
                    bad += hout[s * 1024 + y * 32 + x] != himg[2 * fs + (long long)(33 + y) * pitch + OX + s + x];
            }
            printf("lane_sel %d level %d: ok, mismatches %d\n", lane_sel, level, bad);
        }
    }
    return 0;
}
