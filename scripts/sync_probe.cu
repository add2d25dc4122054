// sync_probe.cu -- what compute-sanitizer's synccheck says about named barriers
// that only a subset of a CTA's warps enter (the pattern of the forward kernel:
// compute warps + scribe warps at barrier 1, everybody at barrier 2, the two
// scribe warps alone at barrier 3).  Every variant is a correct program.
//   nvcc -arch=sm_100a -o sync_probe sync_probe.cu
//   compute-sanitizer --tool synccheck ./sync_probe <variant>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

// variant 0: all four warps at barrier 1 (count 128), then at barrier 2
// variant 1: warps 0-2 at barrier 1 (count 96), all four at barrier 2
// variant 2: as 1, and warps 1-2 alone at barrier 3 (count 64) in between
// variant 3: as 2 with the thread counts in registers (computed from blockDim)
// variant 4: as 3, and the warps do different amounts of work in front of the
//            barriers (shuffles under warp-uniform conditions, like the kernel)
__global__ void probe(int variant, int iters, int *out)
{
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nall = blockDim.x, n1 = blockDim.x - 32;
    int acc = 0;
    for (int i = 0; i < iters; i++) {
        if (variant == 0) {
            bar_sync(1, 128);
        } else if (warp < 3) {
            if (variant >= 4 && warp == 2)
                for (int l = 0; l < 3; l++)
                    if ((1 << l) <= (i & 3))
                        acc += __shfl_up_sync(0xffffffffu, acc + lane, 1 << l);
            if (variant == 5 && warp == 2) {
                // variant 5: the same named barrier entered from a different place
                // in the code (producer and consumer loops of their own)
                asm volatile("// consumer side");
                acc += 7;
                bar_sync(1, n1);
                acc ^= 3;
            } else {
                bar_sync(1, variant >= 3 ? n1 : 96);
            }
            if (variant >= 2 && warp >= 1) {
                if (variant >= 4 && (lane & 1))
                    acc ^= i;
                bar_sync(3, 64);
            }
            if (variant >= 4 && warp == 1 && lane < 19)
                acc += lane * i;
        }
        acc += i;
        bar_sync(2, variant >= 3 ? nall : 128);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        out[blockIdx.x] = acc;
}

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    int *out;
    cudaMalloc(&out, sizeof(int) * 4);
    probe<<<4, 128>>>(variant, 100, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: %s\n", variant, cudaGetErrorString(e));
    return 0;
}
