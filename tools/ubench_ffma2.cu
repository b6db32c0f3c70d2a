// Micro-benchmark: issue rate of FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100) per SM sub-partition, with ILP chains.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_ab/ubench_ffma2 tools/ubench_ffma2.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int ILP, bool PACKED>
__global__ void __launch_bounds__(256, 2) k(int iters, float seed, float* sink) {
	float2 a[ILP];
	const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(seed, -seed);
#pragma unroll
	for (int i = 0; i < ILP; ++i) a[i] = make_float2(seed + i, seed - i);
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int i = 0; i < ILP; ++i) {
				if (PACKED) a[i] = __ffma2_rn(a[i], b, c);
				else {
					a[i].x = fmaf(a[i].x, b.x, c.x);
					a[i].y = fmaf(a[i].y, b.y, c.y);
				}
			}
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y;
	if (s == 12345.678f) *sink = s;
}

template <int ILP, bool PACKED>
void run(const char* name, int warps_per_block) {
	float* sink;
	cudaMalloc(&sink, 4);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int iters = 20000, blocks = 148 * 2;
	k<ILP, PACKED><<<blocks, warps_per_block * 32>>>(100, 1.0f, sink);
	cudaEventRecord(e0);
	k<ILP, PACKED><<<blocks, warps_per_block * 32>>>(iters, 1.0f, sink);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	double fma = (double)blocks * warps_per_block * 32 * iters * 8.0 * ILP * 2;   // scalar FMAs
	printf("%-28s warps/block %d: %.3f ms, %.2f TFLOP/s\n", name, warps_per_block, ms, 2 * fma / ms / 1e9);
}

int main() {
	for (int w : {1, 2, 4, 8}) {
		run<1, false>("FFMA  ilp1(x2 scalars)", w);
		run<1, true>("FFMA2 ilp1", w);
		run<4, false>("FFMA  ilp4(x2 scalars)", w);
		run<4, true>("FFMA2 ilp4", w);
	}
	return 0;
}
