// GPU-side CSR construction from an edge list: the device twin of Graph::setupImporterNew
// (graph/graphCPU.cpp:112-170).  Same result, array for array: self-loops dropped, the back-edge of every edge added,
// duplicates kept, and inside a row the neighbours in FILE ORDER (edge i contributes dst to src's row and src to dst's row,
// in the order the reference's two-pass fill visits them) -- a stable radix sort of the doubled edge list by row does exactly that.
#pragma once
#include <cub/cub.cuh>
#include <cstdint>

namespace mcmcb200 {

// pair 2i = (src -> dst), pair 2i+1 = (dst -> src); self-loops and out-of-range ids get the key n (sorted behind every row)
__global__ void csr_pairs_kernel(const uint32_t * src, const uint32_t * dst, uint64_t m, uint32_t n, uint32_t * keys, uint32_t * vals,
                                 uint32_t * deg /* [n+1], zeroed */, unsigned long long * bad) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const uint32_t s = src[i], d = dst[i];
	const bool inRange = s < n && d < n;
	const bool ok = inRange && s != d;
	if (!inRange) atomicAdd(bad, 1ull);
	keys[2 * i] = ok ? s : n;  vals[2 * i] = d;
	keys[2 * i + 1] = ok ? d : n;  vals[2 * i + 1] = s;
	if (ok) { atomicAdd(&deg[s], 1u); atomicAdd(&deg[d], 1u); }
}

// d_rowptr [n+1] and d_neighs [nnz + 16, zero padded] are cudaMalloc'ed here; src / dst are DEVICE arrays of m entries.
// Returns cudaErrorInvalidValue with *badOut > 0 if an endpoint is >= n.
inline cudaError_t build_csr_from_edges(uint32_t n, uint64_t m, const uint32_t * d_src, const uint32_t * d_dst, cudaStream_t stream,
                                        uint32_t ** rowptrOut, uint32_t ** neighsOut, uint64_t * nnzOut, uint64_t * badOut) {
	*rowptrOut = nullptr; *neighsOut = nullptr; *nnzOut = 0; *badOut = 0;
	if (2 * m >= 0xfffffff0ull) return cudaErrorInvalidValue;          // 32-bit CSR offsets, like the reference (graph.h:19-20)
	if (n >= 0x7ffffff0u) return cudaErrorInvalidValue;                 // cub::DeviceScan item counts are int
	cudaError_t err = cudaSuccess;
	const uint64_t pairs = 2 * m;
	uint32_t * d_keys[2] = {nullptr, nullptr}, * d_vals[2] = {nullptr, nullptr}, * d_deg = nullptr, * d_rowptr = nullptr, * d_neighs = nullptr;
	unsigned long long * d_bad = nullptr;
	void * d_tmp = nullptr; size_t tmpBytes = 0, tmp2 = 0;
	unsigned long long bad = 0; uint32_t nnz = 0;
	int endBit = 1;
	while (endBit < 32 && (1ull << endBit) <= (unsigned long long)n) endBit++;   // keys are in [0, n]
#define CSR_CU(call) do { err = (call); if (err != cudaSuccess) goto done; } while (0)
	for (int i = 0; i < 2; ++i) {
		CSR_CU(cudaMalloc(&d_keys[i], sizeof(uint32_t) * std::max<uint64_t>(pairs, 1)));
		CSR_CU(cudaMalloc(&d_vals[i], sizeof(uint32_t) * std::max<uint64_t>(pairs, 1)));
	}
	CSR_CU(cudaMalloc(&d_deg, sizeof(uint32_t) * ((size_t)n + 2)));
	CSR_CU(cudaMalloc(&d_rowptr, sizeof(uint32_t) * ((size_t)n + 2)));
	CSR_CU(cudaMalloc(&d_bad, sizeof(unsigned long long)));
	CSR_CU(cudaMemsetAsync(d_deg, 0, sizeof(uint32_t) * ((size_t)n + 2), stream));
	CSR_CU(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), stream));
	if (m) csr_pairs_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(d_src, d_dst, m, n, d_keys[0], d_vals[0], d_deg, d_bad);
	CSR_CU(cudaGetLastError());
	{
		cub::DoubleBuffer<uint32_t> kb(d_keys[0], d_keys[1]), vb(d_vals[0], d_vals[1]);
		CSR_CU(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, kb, vb, (long long)pairs, 0, endBit, stream));
		CSR_CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_deg, d_rowptr, (int)n + 1, stream));
		tmpBytes = std::max(tmpBytes, tmp2);
		CSR_CU(cudaMalloc(&d_tmp, std::max<size_t>(tmpBytes, 16)));
		if (pairs) CSR_CU(cub::DeviceRadixSort::SortPairs(d_tmp, tmpBytes, kb, vb, (long long)pairs, 0, endBit, stream));   // stable: file order inside a row
		CSR_CU(cub::DeviceScan::ExclusiveSum(d_tmp, tmp2, d_deg, d_rowptr, (int)n + 1, stream));
		CSR_CU(cudaMemcpyAsync(&nnz, d_rowptr + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
		CSR_CU(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, stream));
		CSR_CU(cudaStreamSynchronize(stream));
		if (bad) { *badOut = bad; err = cudaErrorInvalidValue; goto done; }
		CSR_CU(cudaMalloc(&d_neighs, sizeof(uint32_t) * ((size_t)nnz + 16)));
		CSR_CU(cudaMemsetAsync(d_neighs, 0, sizeof(uint32_t) * ((size_t)nnz + 16), stream));
		if (nnz) CSR_CU(cudaMemcpyAsync(d_neighs, vb.Current(), sizeof(uint32_t) * (size_t)nnz, cudaMemcpyDeviceToDevice, stream));
		CSR_CU(cudaStreamSynchronize(stream));
	}
	*rowptrOut = d_rowptr; *neighsOut = d_neighs; *nnzOut = nnz;
	d_rowptr = nullptr; d_neighs = nullptr;
done:
#undef CSR_CU
	cudaFree(d_keys[0]); cudaFree(d_keys[1]); cudaFree(d_vals[0]); cudaFree(d_vals[1]); cudaFree(d_deg); cudaFree(d_bad); cudaFree(d_tmp);
	cudaFree(d_rowptr); cudaFree(d_neighs);
	return err;
}

} // namespace mcmcb200
