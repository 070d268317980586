/* The reference-side binding of libcstone_b200.so: definitions of the `extern template` GPU entry points that the
 * reference declares in the _gpu.h headers under include/cstone and normally implements in its static library cstone_gpu
 * (include/cstone/CMakeLists.txt:9-22), here as forwarders to the C ABI of this repository (include/cstone_b200.h).
 *
 * This file is compiled AGAINST THE REFERENCE'S OWN HEADERS (-I <reference>/include): every definition below has the
 * reference's exact signature and is explicitly instantiated for the types the reference instantiates, so the
 * header-only layers of the reference and its tests link against it unchanged.  It contains no algorithm: element-wise
 * helpers of primitives_gpu.h that have no counterpart in the C ABI (fill, sequence, generic gather / scatter) are a few
 * lines of CUDA each; everything on the hot path forwards.
 *
 * Covered (file:line = the declaration in the reference):
 *   cuda/device_vector.h:32-94                DeviceVector<T> (15 element types) and operator==
 *   sfc/sfc_gpu.h:24-26                       computeSfcKeys, {Morton,Hilbert}Key<{unsigned,uint64_t}> x {float,double}
 *   primitives/primitives_gpu.h:30-174        fill, sequence, gather, scatter, sort, sortByKey, sortByKeyTempStorage,
 *                                             exclusiveScan, lowerBound (both forms)
 *   tree/csarray_gpu.h:41-82                  computeNodeCountsGpu, computeNodeOpsGpu, rebalanceTreeGpu,
 *                                             countSfcGapsGpu, fillSfcGapsGpu
 *   tree/octree_gpu.h:35-71                   buildOctreeGpu (both overloads), upsweepSumGpu
 *   traversal/collisions_gpu.h:46-58          findHalosGpu
 *   focus/source_center_gpu.h:40-52,100-106   computeBoundingBoxGpu, computeGeoCentersGpu
 * Not covered yet (the Domain-level binding goes through cs_domain_*): markMacsGpu, the LET
 * rebalance entry points, gatherRanges, groups, the remaining primitives.
 */
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "cstone/cuda/device_vector.h"
#include "cstone/focus/source_center_gpu.h"
#include "cstone/primitives/primitives_gpu.h"
#include "cstone/sfc/sfc_gpu.h"
#include "cstone/traversal/groups_gpu.h"
#include "cstone/halos/gather_halos_gpu.h"
#include "cstone/traversal/collisions_gpu.h"
#include "cstone/tree/csarray_gpu.h"
#include "cstone/tree/octree_gpu.h"

#include "cstone_b200.h"

namespace cstone
{

namespace
{

//! status -> the reference's two error behaviours: CUDA failures print and exit (cuda/errorcheck.cuh:15-27),
//! contract violations throw std::runtime_error (primitives/primitives_gpu.cu:338)
void csCheck(int status, const char* what)
{
    if (status == 0) { return; }
    if (status == 1)
    {
        std::fprintf(stderr, "%s: CUDA error: %s\n", what, cs_last_error());
        std::exit(EXIT_FAILURE);
    }
    throw std::runtime_error(std::string(what) + ": " + cs_last_error());
}

void cudaCheck(cudaError_t e, const char* what)
{
    if (e != cudaSuccess)
    {
        std::fprintf(stderr, "%s: CUDA error: %s\n", what, cudaGetErrorString(e));
        std::exit(EXIT_FAILURE);
    }
}

template<class T>
struct BoxArgs
{
    double lim[6];
    int bnd[3];
    explicit BoxArgs(const Box<T>& b)
        : lim{double(b.xmin()), double(b.xmax()), double(b.ymin()), double(b.ymax()), double(b.zmin()), double(b.zmax())}
        , bnd{int(b.boundaryX()), int(b.boundaryY()), int(b.boundaryZ())}
    {
    }
};

//! grow-only device scratch per stream for entry points whose C-ABI counterpart takes caller-provided temporaries
void* scratch(cudaStream_t s, int slot, size_t bytes)
{
    static std::mutex m;
    static std::map<std::pair<cudaStream_t, int>, std::pair<void*, size_t>> bufs;
    std::lock_guard<std::mutex> lk(m);
    auto& b = bufs[{s, slot}];
    if (bytes > b.second)
    {
        if (b.first)
        {
            cudaStreamSynchronize(s);
            cudaFree(b.first);
        }
        cudaCheck(cudaMalloc(&b.first, bytes + bytes / 4 + 256), "scratch");
        b.second = bytes + bytes / 4 + 256;
    }
    return b.first;
}

template<class T>
__global__ void fillKernel(T* p, size_t n, T v)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { p[i] = v; }
}

template<class I>
__global__ void sequenceKernel(I* p, size_t n, I init)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { p[i] = init + I(i); }
}

template<class TS, class TD, class I>
__global__ void gatherKernel(const I* ord, size_t n, const TS* src, TD* dst)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { dst[i] = TD(src[ord[i]]); }
}

template<class T, class I>
__global__ void scatterKernel(const I* ord, size_t n, const T* src, T* dst)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { dst[ord[i]] = src[i]; }
}

template<class T>
__global__ void equalKernel(const T* a, const T* b, size_t n, int* differ)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
    {
        const unsigned char* pa = reinterpret_cast<const unsigned char*>(a + i);
        const unsigned char* pb = reinterpret_cast<const unsigned char*>(b + i);
        for (size_t k = 0; k < sizeof(T); ++k)
            if (pa[k] != pb[k]) { *differ = 1; }
    }
}

template<class T, class I>
__global__ void lowerBoundKernel(const T* first, size_t n, const T* values, size_t numValues, I* result)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= numValues) { return; }
    T v       = values[i];
    size_t lo = 0, hi = n;
    while (lo < hi)
    {
        size_t m = lo + (hi - lo) / 2;
        if (first[m] < v) { lo = m + 1; }
        else { hi = m; }
    }
    result[i] = I(lo);
}

template<class T>
__global__ void lowerBoundScalarKernel(const T* first, size_t n, T v, unsigned long long* result)
{
    size_t lo = 0, hi = n;
    while (lo < hi)
    {
        size_t m = lo + (hi - lo) / 2;
        if (first[m] < v) { lo = m + 1; }
        else { hi = m; }
    }
    *result = lo;
}

unsigned blocks(size_t n) { return unsigned((n + 255) / 256); }

} // namespace

/* ------------------------------------------------------------------------------------------------ DeviceVector */

template<class T>
class DeviceVector<T>::Impl
{
public:
    T* p{nullptr};
    std::size_t n{0}, cap{0};

    ~Impl() { cudaFree(p); }

    void reserve(std::size_t c)
    {
        if (c <= cap) { return; }
        T* q = nullptr;
        cudaCheck(cudaMalloc(&q, c * sizeof(T)), "DeviceVector");
        if (n) { cudaCheck(cudaMemcpy(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice), "DeviceVector"); }
        cudaFree(p);
        p   = q;
        cap = c;
    }
    void resize(std::size_t m)
    {
        reserve(m);
        // new elements are value-initialised like thrust::device_vector::resize
        if (m > n) { cudaCheck(cudaMemset(p + n, 0, (m - n) * sizeof(T)), "DeviceVector"); }
        n = m;
    }
    void shrink()
    {
        if (cap == n) { return; }
        T* q = nullptr;
        if (n)
        {
            cudaCheck(cudaMalloc(&q, n * sizeof(T)), "DeviceVector");
            cudaCheck(cudaMemcpy(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice), "DeviceVector");
        }
        cudaFree(p);
        p   = q;
        cap = n;
    }
};

template<class T>
DeviceVector<T>::DeviceVector()
    : impl_(new Impl())
{
}
template<class T>
DeviceVector<T>::DeviceVector(std::size_t size)
    : impl_(new Impl())
{
    impl_->resize(size);
}
template<class T>
DeviceVector<T>::DeviceVector(std::size_t size, T init)
    : impl_(new Impl())
{
    impl_->reserve(size);
    impl_->n = size;
    if (size)
    {
        std::vector<T> h(size, init);
        cudaCheck(cudaMemcpy(impl_->p, h.data(), size * sizeof(T), cudaMemcpyHostToDevice), "DeviceVector");
    }
}
template<class T>
DeviceVector<T>::DeviceVector(const DeviceVector<T>& other)
    : impl_(new Impl())
{
    impl_->reserve(other.size());
    impl_->n = other.size();
    if (other.size())
    {
        cudaCheck(cudaMemcpy(impl_->p, other.data(), other.size() * sizeof(T), cudaMemcpyDeviceToDevice), "DeviceVector");
    }
}
template<class T>
DeviceVector<T>::DeviceVector(const std::vector<T>& rhs)
    : impl_(new Impl())
{
    *this = rhs;
}
template<class T>
DeviceVector<T>::DeviceVector(const T* first, const T* last)
    : impl_(new Impl())
{
    std::size_t size = last - first;
    impl_->reserve(size);
    impl_->n = size;
    if (size) { cudaCheck(cudaMemcpy(impl_->p, first, size * sizeof(T), cudaMemcpyHostToDevice), "DeviceVector"); }
}
template<class T>
DeviceVector<T>::~DeviceVector() = default;

template<class T>
T* DeviceVector<T>::data()
{
    return impl_->p;
}
template<class T>
const T* DeviceVector<T>::data() const
{
    return impl_->p;
}
template<class T>
T* DeviceVector<T>::begin()
{
    return impl_->p;
}
template<class T>
const T* DeviceVector<T>::cbegin() const
{
    return impl_->p;
}
template<class T>
T* DeviceVector<T>::end()
{
    return impl_->p + impl_->n;
}
template<class T>
const T* DeviceVector<T>::cend() const
{
    return impl_->p + impl_->n;
}
template<class T>
void DeviceVector<T>::resize(std::size_t size)
{
    impl_->resize(size);
}
template<class T>
void DeviceVector<T>::reserve(std::size_t size)
{
    impl_->reserve(size);
}
template<class T>
void DeviceVector<T>::shrink_to_fit()
{
    impl_->shrink();
}
template<class T>
std::size_t DeviceVector<T>::size() const
{
    return impl_->n;
}
template<class T>
bool DeviceVector<T>::empty() const
{
    return impl_->n == 0;
}
template<class T>
std::size_t DeviceVector<T>::capacity() const
{
    return impl_->cap;
}
template<class T>
DeviceVector<T>& DeviceVector<T>::swap(DeviceVector<T>& rhs)
{
    std::swap(impl_, rhs.impl_);
    return *this;
}
template<class T>
DeviceVector<T>& DeviceVector<T>::operator=(const std::vector<T>& rhs)
{
    impl_->reserve(rhs.size());
    impl_->n = rhs.size();
    if (rhs.size())
    {
        cudaCheck(cudaMemcpy(impl_->p, rhs.data(), rhs.size() * sizeof(T), cudaMemcpyHostToDevice), "DeviceVector");
    }
    return *this;
}
template<class T>
DeviceVector<T>& DeviceVector<T>::operator=(DeviceVector<T> rhs)
{
    swap(rhs);
    return *this;
}

template<class T>
bool operator==(const DeviceVector<T>& lhs, const DeviceVector<T>& rhs)
{
    if (lhs.size() != rhs.size()) { return false; }
    if (lhs.size() == 0) { return true; }
    int* differ = nullptr;
    cudaCheck(cudaMalloc(&differ, sizeof(int)), "DeviceVector ==");
    cudaCheck(cudaMemset(differ, 0, sizeof(int)), "DeviceVector ==");
    equalKernel<<<blocks(lhs.size()), 256>>>(lhs.data(), rhs.data(), lhs.size(), differ);
    int h = 0;
    cudaCheck(cudaMemcpy(&h, differ, sizeof(int), cudaMemcpyDeviceToHost), "DeviceVector ==");
    cudaFree(differ);
    return h == 0;
}

#define CS_DEVICE_VECTOR(T)                                                                                            \
    template class DeviceVector<T>;                                                                                    \
    template bool operator==(const DeviceVector<T>&, const DeviceVector<T>&);

CS_DEVICE_VECTOR(char)
CS_DEVICE_VECTOR(uint8_t)
CS_DEVICE_VECTOR(int)
CS_DEVICE_VECTOR(unsigned)
CS_DEVICE_VECTOR(uint64_t)
CS_DEVICE_VECTOR(float)
CS_DEVICE_VECTOR(double)
using ArrI2 = util::array<int, 2>;
using ArrI3 = util::array<int, 3>;
using ArrU1 = util::array<unsigned, 1>;
using ArrU2 = util::array<unsigned, 2>;
using ArrF3 = util::array<float, 3>;
using ArrD3 = util::array<double, 3>;
using ArrF4 = util::array<float, 4>;
using ArrD4 = util::array<double, 4>;
CS_DEVICE_VECTOR(ArrI2)
CS_DEVICE_VECTOR(ArrI3)
CS_DEVICE_VECTOR(ArrU1)
CS_DEVICE_VECTOR(ArrU2)
CS_DEVICE_VECTOR(ArrF3)
CS_DEVICE_VECTOR(ArrD3)
CS_DEVICE_VECTOR(ArrF4)
CS_DEVICE_VECTOR(ArrD4)

/* ------------------------------------------------------------------------------------------------ sfc_gpu.h */

template<class KeyType, class T>
void computeSfcKeys(execution::Gpu exec, const T* x, const T* y, const T* z, KeyType* keys, size_t numKeys,
                    const Box<T>& box)
{
    using Integer  = typename KeyType::ValueType;
    const int kind = IsMorton<KeyType>{} ? 1 : 0;
    BoxArgs<T> b(box);
    if constexpr (sizeof(Integer) == 4 && std::is_same_v<T, float>)
    {
        csCheck(cs_compute_sfc_keys_u32f(kind, x, y, z, reinterpret_cast<uint32_t*>(keys), numKeys, b.lim, b.bnd,
                                         cudaStream_t(exec)),
                "computeSfcKeys");
    }
    else if constexpr (sizeof(Integer) == 4)
    {
        csCheck(cs_compute_sfc_keys_u32d(kind, x, y, z, reinterpret_cast<uint32_t*>(keys), numKeys, b.lim, b.bnd,
                                         cudaStream_t(exec)),
                "computeSfcKeys");
    }
    else if constexpr (std::is_same_v<T, float>)
    {
        csCheck(cs_compute_sfc_keys_u64f(kind, x, y, z, reinterpret_cast<uint64_t*>(keys), numKeys, b.lim, b.bnd,
                                         cudaStream_t(exec)),
                "computeSfcKeys");
    }
    else
    {
        csCheck(cs_compute_sfc_keys_u64d(kind, x, y, z, reinterpret_cast<uint64_t*>(keys), numKeys, b.lim, b.bnd,
                                         cudaStream_t(exec)),
                "computeSfcKeys");
    }
}

#define CS_SFC_KEYS(KeyType, T)                                                                                        \
    template void computeSfcKeys(execution::Gpu, const T*, const T*, const T*, KeyType*, size_t, const Box<T>&);
CS_SFC_KEYS(MortonKey<unsigned>, float)
CS_SFC_KEYS(MortonKey<unsigned>, double)
CS_SFC_KEYS(MortonKey<uint64_t>, float)
CS_SFC_KEYS(MortonKey<uint64_t>, double)
CS_SFC_KEYS(HilbertKey<unsigned>, float)
CS_SFC_KEYS(HilbertKey<unsigned>, double)
CS_SFC_KEYS(HilbertKey<uint64_t>, float)
CS_SFC_KEYS(HilbertKey<uint64_t>, double)

/* ------------------------------------------------------------------------------------------------ primitives_gpu.h */

template<class T>
void fill(execution::Gpu exec, T* first, T* last, T value)
{
    size_t n = last - first;
    if (n) { fillKernel<<<blocks(n), 256, 0, exec>>>(first, n, value); }
}
template void fill(execution::Gpu, double*, double*, double);
template void fill(execution::Gpu, float*, float*, float);
template void fill(execution::Gpu, int*, int*, int);
template void fill(execution::Gpu, uint8_t*, uint8_t*, uint8_t);
template void fill(execution::Gpu, char*, char*, char);
template void fill(execution::Gpu, unsigned*, unsigned*, unsigned);
template void fill(execution::Gpu, uint64_t*, uint64_t*, uint64_t);

template<class IndexType>
void sequence(execution::Gpu exec, IndexType* input, size_t numElements, IndexType init)
{
    if constexpr (sizeof(IndexType) == 4)
    {
        csCheck(cs_sequence_u32(uint32_t(init), numElements, reinterpret_cast<uint32_t*>(input), cudaStream_t(exec)),
                "sequence");
    }
    else if (numElements) { sequenceKernel<<<blocks(numElements), 256, 0, exec>>>(input, numElements, init); }
}
template void sequence(execution::Gpu, int*, size_t, int);
template void sequence(execution::Gpu, unsigned*, size_t, unsigned);
template void sequence(execution::Gpu, uint64_t*, uint64_t, uint64_t);

template<class TS, class TD, class IndexType>
void gather(execution::Gpu exec, const IndexType* ordering, size_t numElements, const TS* src, TD* buffer)
{
    if (numElements == 0) { return; }
    if constexpr (std::is_same_v<TS, TD> && sizeof(IndexType) == 4 && (sizeof(TS) == 4 || sizeof(TS) == 8))
    {
        csCheck(cs_gather(reinterpret_cast<const uint32_t*>(ordering), numElements, src, buffer, int(sizeof(TS)),
                          cudaStream_t(exec)),
                "gather");
    }
    else { gatherKernel<<<blocks(numElements), 256, 0, exec>>>(ordering, numElements, src, buffer); }
}
#define CS_GATHER(I, TS, TD) template void gather(execution::Gpu, const I*, size_t, const TS*, TD*);
using ArrF1  = util::array<float, 1>;
using ArrF2  = util::array<float, 2>;
using ArrF8  = util::array<float, 8>;
using ArrF12 = util::array<float, 12>;
using ArrD8  = util::array<double, 8>;
using ArrD12 = util::array<double, 12>;
CS_GATHER(int, uint8_t, uint32_t)
CS_GATHER(int, int, int)
CS_GATHER(int, uint32_t, uint32_t)
CS_GATHER(int, uint64_t, uint64_t)
CS_GATHER(int, ArrF3, ArrF3)
CS_GATHER(int, ArrF4, ArrF4)
CS_GATHER(int, ArrF8, ArrF8)
CS_GATHER(int, ArrF12, ArrF12)
CS_GATHER(int, ArrD3, ArrD3)
CS_GATHER(int, ArrD4, ArrD4)
CS_GATHER(int, ArrD8, ArrD8)
CS_GATHER(int, ArrD12, ArrD12)
CS_GATHER(unsigned, uint8_t, uint8_t)
CS_GATHER(unsigned, double, double)
CS_GATHER(unsigned, float, float)
CS_GATHER(unsigned, char, char)
CS_GATHER(unsigned, int, int)
CS_GATHER(unsigned, long, long)
CS_GATHER(unsigned, unsigned, unsigned)
CS_GATHER(unsigned, unsigned long, unsigned long)
CS_GATHER(unsigned, unsigned long long, unsigned long long)
CS_GATHER(unsigned, ArrF1, ArrF1)
CS_GATHER(unsigned, ArrF2, ArrF2)
CS_GATHER(unsigned, ArrF3, ArrF3)
CS_GATHER(unsigned, ArrF4, ArrF4)

template<class T, class IndexType>
void scatter(execution::Gpu exec, const IndexType* ordering, size_t numElements, const T* src, T* buffer)
{
    if (numElements) { scatterKernel<<<blocks(numElements), 256, 0, exec>>>(ordering, numElements, src, buffer); }
}
#define CS_SCATTER(T) template void scatter(execution::Gpu, const int*, size_t, const T*, T*);
CS_SCATTER(int)
CS_SCATTER(uint32_t)
CS_SCATTER(uint64_t)
CS_SCATTER(ArrF4)
CS_SCATTER(ArrF8)
CS_SCATTER(ArrF12)
CS_SCATTER(ArrD4)
CS_SCATTER(ArrD8)
CS_SCATTER(ArrD12)

template<class T, class IndexType>
void lowerBound(execution::Gpu exec, const T* first, const T* last, const T* valueFirst, const T* valueLast,
                IndexType* result)
{
    size_t nv = valueLast - valueFirst;
    if (nv) { lowerBoundKernel<<<blocks(nv), 256, 0, exec>>>(first, size_t(last - first), valueFirst, nv, result); }
}
template<class T>
size_t lowerBound(execution::Gpu exec, const T* first, const T* last, T value)
{
    auto* d = static_cast<unsigned long long*>(scratch(exec, 4, sizeof(unsigned long long)));
    lowerBoundScalarKernel<<<1, 1, 0, exec>>>(first, size_t(last - first), value, d);
    unsigned long long h = 0;
    cudaCheck(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, exec), "lowerBound");
    cudaCheck(cudaStreamSynchronize(exec), "lowerBound");
    return size_t(h);
}
template size_t lowerBound(execution::Gpu, const unsigned*, const unsigned*, unsigned);
template size_t lowerBound(execution::Gpu, const uint64_t*, const uint64_t*, uint64_t);
template size_t lowerBound(execution::Gpu, const int*, const int*, int);
template size_t lowerBound(execution::Gpu, const int64_t*, const int64_t*, int64_t);
template size_t lowerBound(execution::Gpu, const float*, const float*, float);

template void lowerBound(execution::Gpu, const unsigned*, const unsigned*, const unsigned*, const unsigned*, unsigned*);
template void lowerBound(execution::Gpu, const uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*, unsigned*);
template void lowerBound(execution::Gpu, const unsigned*, const unsigned*, const unsigned*, const unsigned*, uint64_t*);
template void lowerBound(execution::Gpu, const uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*, uint64_t*);

template<class KeyType>
void sort(execution::Gpu exec, KeyType* first, KeyType* last, KeyType* keyBuf)
{
    static_assert(sizeof(KeyType) == 4 || sizeof(KeyType) == 8);
    size_t n = last - first;
    if constexpr (sizeof(KeyType) == 8)
    {
        size_t tb = cs_sort_by_key_temp_bytes_u64(n);
        csCheck(cs_sort_by_key_u64(reinterpret_cast<uint64_t*>(first), nullptr, n, reinterpret_cast<uint64_t*>(keyBuf),
                                   nullptr, scratch(exec, 0, tb), tb, cudaStream_t(exec)),
                "sort");
    }
    else
    {
        size_t tb = cs_sort_by_key_temp_bytes_u32(n);
        csCheck(cs_sort_by_key_u32(reinterpret_cast<uint32_t*>(first), nullptr, n, reinterpret_cast<uint32_t*>(keyBuf),
                                   nullptr, scratch(exec, 0, tb), tb, cudaStream_t(exec)),
                "sort");
    }
}
template void sort(execution::Gpu, uint32_t*, uint32_t*, uint32_t*);
template void sort(execution::Gpu, uint64_t*, uint64_t*, uint64_t*);

template<class KeyType, class ValueType>
uint64_t sortByKeyTempStorage(uint64_t numElements)
{
    return sizeof(KeyType) == 8 ? cs_sort_by_key_temp_bytes_u64(numElements) : cs_sort_by_key_temp_bytes_u32(numElements);
}

template<class KeyType, class ValueType>
void sortByKey(execution::Gpu exec, KeyType* first, KeyType* last, ValueType* values, KeyType* keyBuf,
               ValueType* valueBuf, void* tmp, uint64_t tmpBytes)
{
    static_assert(sizeof(ValueType) == 4, "values are LocalIndex / TreeNodeIndex");
    size_t n = last - first;
    // "temp storage too small" is a contract violation in the reference as well (primitives_gpu.cu:338): status 2 throws
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_sort_by_key_u64(reinterpret_cast<uint64_t*>(first), reinterpret_cast<uint32_t*>(values), n,
                                   reinterpret_cast<uint64_t*>(keyBuf), reinterpret_cast<uint32_t*>(valueBuf), tmp,
                                   tmpBytes, cudaStream_t(exec)),
                "sortByKey");
    }
    else
    {
        csCheck(cs_sort_by_key_u32(reinterpret_cast<uint32_t*>(first), reinterpret_cast<uint32_t*>(values), n,
                                   reinterpret_cast<uint32_t*>(keyBuf), reinterpret_cast<uint32_t*>(valueBuf), tmp,
                                   tmpBytes, cudaStream_t(exec)),
                "sortByKey");
    }
}
#define CS_SORT_BY_KEY(KeyType, ValueType)                                                                             \
    template void sortByKey(execution::Gpu, KeyType*, KeyType*, ValueType*, KeyType*, ValueType*, void*, uint64_t);    \
    template uint64_t sortByKeyTempStorage<KeyType, ValueType>(uint64_t);
CS_SORT_BY_KEY(uint32_t, uint32_t)
CS_SORT_BY_KEY(uint32_t, int)
CS_SORT_BY_KEY(uint64_t, uint32_t)
CS_SORT_BY_KEY(uint64_t, int)

template<class IndexType, class SumType>
void exclusiveScan(execution::Gpu exec, const IndexType* first, const IndexType* last, SumType* output, SumType init)
{
    static_assert(sizeof(IndexType) == 4 && sizeof(SumType) == 4, "32-bit scans");
    size_t n = last - first;
    if (n == 0) { return; }
    if (init != SumType(0)) { throw std::runtime_error("exclusiveScan: non-zero init is not supported"); }
    size_t tb = cs_scan_temp_bytes(n);
    csCheck(cs_exclusive_scan_u32(reinterpret_cast<const uint32_t*>(first), reinterpret_cast<uint32_t*>(output), n,
                                  scratch(exec, 1, tb), cudaStream_t(exec)),
            "exclusiveScan");
}
template void exclusiveScan(execution::Gpu, const int*, const int*, int*, int);
template void exclusiveScan(execution::Gpu, const int*, const int*, unsigned*, unsigned);
template void exclusiveScan(execution::Gpu, const unsigned*, const unsigned*, unsigned*, unsigned);

/* ------------------------------------------------------------------------------------------------ csarray_gpu.h */

template<class KeyType>
void computeNodeCountsGpu(execution::Gpu exec, const KeyType* tree, unsigned* counts, TreeNodeIndex numNodes,
                          std::span<const KeyType> keys, unsigned maxCount, bool /*useCountsAsGuess*/)
{
    // the guesses only narrow the reference's binary searches; the counts they produce are the same
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_compute_node_counts_u64(reinterpret_cast<const uint64_t*>(tree), counts, numNodes,
                                           reinterpret_cast<const uint64_t*>(keys.data()), keys.size(), maxCount,
                                           cudaStream_t(exec)),
                "computeNodeCountsGpu");
    }
    else
    {
        csCheck(cs_compute_node_counts_u32(reinterpret_cast<const uint32_t*>(tree), counts, numNodes,
                                           reinterpret_cast<const uint32_t*>(keys.data()), keys.size(), maxCount,
                                           cudaStream_t(exec)),
                "computeNodeCountsGpu");
    }
}
template void computeNodeCountsGpu(execution::Gpu, const unsigned*, unsigned*, TreeNodeIndex, std::span<const unsigned>,
                                   unsigned, bool);
template void computeNodeCountsGpu(execution::Gpu, const uint64_t*, unsigned*, TreeNodeIndex, std::span<const uint64_t>,
                                   unsigned, bool);

namespace
{
//! the reference keeps "did any node change" in a __device__ global between computeNodeOpsGpu and rebalanceTreeGpu
//! (csarray_gpu.cu:130,226-230); here it travels per stream on the host
std::mutex convergedMutex;
std::map<cudaStream_t, int> convergedOf;
} // namespace

template<class KeyType>
TreeNodeIndex computeNodeOpsGpu(execution::Gpu exec, const KeyType* tree, TreeNodeIndex numNodes, const unsigned* counts,
                                unsigned bucketSize, TreeNodeIndex* nodeOps)
{
    int newNumNodes = 0, converged = 0;
    void* tmp = scratch(exec, 2, cs_node_ops_temp_bytes(numNodes));
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_compute_node_ops_u64(reinterpret_cast<const uint64_t*>(tree), numNodes, counts, bucketSize, nodeOps,
                                        tmp, &newNumNodes, &converged, cudaStream_t(exec)),
                "computeNodeOpsGpu");
    }
    else
    {
        csCheck(cs_compute_node_ops_u32(reinterpret_cast<const uint32_t*>(tree), numNodes, counts, bucketSize, nodeOps,
                                        tmp, &newNumNodes, &converged, cudaStream_t(exec)),
                "computeNodeOpsGpu");
    }
    std::lock_guard<std::mutex> lk(convergedMutex);
    convergedOf[cudaStream_t(exec)] = converged;
    return newNumNodes;
}
template TreeNodeIndex computeNodeOpsGpu(execution::Gpu, const unsigned*, TreeNodeIndex, const unsigned*, unsigned,
                                         TreeNodeIndex*);
template TreeNodeIndex computeNodeOpsGpu(execution::Gpu, const uint64_t*, TreeNodeIndex, const unsigned*, unsigned,
                                         TreeNodeIndex*);

template<class KeyType>
bool rebalanceTreeGpu(execution::Gpu exec, const KeyType* tree, TreeNodeIndex numNodes, TreeNodeIndex newNumNodes,
                      const TreeNodeIndex* nodeOps, KeyType* newTree)
{
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_rebalance_tree_u64(reinterpret_cast<const uint64_t*>(tree), numNodes, newNumNodes, nodeOps,
                                      reinterpret_cast<uint64_t*>(newTree), cudaStream_t(exec)),
                "rebalanceTreeGpu");
    }
    else
    {
        csCheck(cs_rebalance_tree_u32(reinterpret_cast<const uint32_t*>(tree), numNodes, newNumNodes, nodeOps,
                                      reinterpret_cast<uint32_t*>(newTree), cudaStream_t(exec)),
                "rebalanceTreeGpu");
    }
    cudaCheck(cudaStreamSynchronize(exec), "rebalanceTreeGpu"); // the reference synchronises here as well
    std::lock_guard<std::mutex> lk(convergedMutex);
    return convergedOf[cudaStream_t(exec)] != 0;
}
template bool rebalanceTreeGpu(execution::Gpu, const unsigned*, TreeNodeIndex, TreeNodeIndex, const TreeNodeIndex*,
                               unsigned*);
template bool rebalanceTreeGpu(execution::Gpu, const uint64_t*, TreeNodeIndex, TreeNodeIndex, const TreeNodeIndex*,
                               uint64_t*);

template<class KeyType>
void countSfcGapsGpu(execution::Gpu exec, const KeyType* tree, TreeNodeIndex numNodes, TreeNodeIndex* nodeOps)
{
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_count_sfc_gaps_u64(reinterpret_cast<const uint64_t*>(tree), numNodes, nodeOps, cudaStream_t(exec)),
                "countSfcGapsGpu");
    }
    else
    {
        csCheck(cs_count_sfc_gaps_u32(reinterpret_cast<const uint32_t*>(tree), numNodes, nodeOps, cudaStream_t(exec)),
                "countSfcGapsGpu");
    }
}
template void countSfcGapsGpu(execution::Gpu, const uint32_t*, TreeNodeIndex, TreeNodeIndex*);
template void countSfcGapsGpu(execution::Gpu, const uint64_t*, TreeNodeIndex, TreeNodeIndex*);

template<class KeyType>
void fillSfcGapsGpu(execution::Gpu exec, const KeyType* tree, TreeNodeIndex numNodes, const TreeNodeIndex* nodeOps,
                    KeyType* newTree)
{
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_fill_sfc_gaps_u64(reinterpret_cast<const uint64_t*>(tree), numNodes, nodeOps,
                                     reinterpret_cast<uint64_t*>(newTree), cudaStream_t(exec)),
                "fillSfcGapsGpu");
    }
    else
    {
        csCheck(cs_fill_sfc_gaps_u32(reinterpret_cast<const uint32_t*>(tree), numNodes, nodeOps,
                                     reinterpret_cast<uint32_t*>(newTree), cudaStream_t(exec)),
                "fillSfcGapsGpu");
    }
}
template void fillSfcGapsGpu(execution::Gpu, const uint32_t*, TreeNodeIndex, const TreeNodeIndex*, uint32_t*);
template void fillSfcGapsGpu(execution::Gpu, const uint64_t*, TreeNodeIndex, const TreeNodeIndex*, uint64_t*);

/* ------------------------------------------------------------------------------------------------ octree_gpu.h */

namespace
{
template<class KeyType>
void buildOctreeImpl(cudaStream_t s, const KeyType* cstoneTree, OctreeView<KeyType> d, void* tmp, size_t tmpBytes)
{
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_build_octree_u64(reinterpret_cast<const uint64_t*>(cstoneTree), d.numLeafNodes,
                                    reinterpret_cast<uint64_t*>(d.prefixes), d.childOffsets, d.parents, d.d_levelRange,
                                    d.internalToLeaf, d.leafToInternal, tmp, tmpBytes, s),
                "buildOctreeGpu");
    }
    else
    {
        csCheck(cs_build_octree_u32(reinterpret_cast<const uint32_t*>(cstoneTree), d.numLeafNodes,
                                    reinterpret_cast<uint32_t*>(d.prefixes), d.childOffsets, d.parents, d.d_levelRange,
                                    d.internalToLeaf, d.leafToInternal, tmp, tmpBytes, s),
                "buildOctreeGpu");
    }
    // the reference leaves a host copy of the level ranges in the view (octree_gpu.cu:165-167)
    cudaCheck(cudaMemcpyAsync(d.levelRange, d.d_levelRange, (maxTreeLevel<KeyType>{} + 2) * sizeof(TreeNodeIndex),
                              cudaMemcpyDeviceToHost, s),
              "buildOctreeGpu");
    cudaCheck(cudaStreamSynchronize(s), "buildOctreeGpu");
}
} // namespace

template<class KeyType>
void buildOctreeGpu(execution::Gpu exec, const KeyType* cstoneTree, OctreeView<KeyType> d)
{
    size_t tb = sizeof(KeyType) == 8 ? cs_build_octree_temp_bytes_u64(d.numLeafNodes)
                                     : cs_build_octree_temp_bytes_u32(d.numLeafNodes);
    buildOctreeImpl(exec, cstoneTree, d, scratch(exec, 3, tb), tb);
}
template void buildOctreeGpu(execution::Gpu, const uint32_t*, OctreeView<uint32_t>);
template void buildOctreeGpu(execution::Gpu, const uint64_t*, OctreeView<uint64_t>);

template<class KeyType>
void buildOctreeGpu(execution::Gpu exec, const KeyType* cstoneTree, OctreeView<KeyType> d, std::span<KeyType>,
                    std::span<TreeNodeIndex>, std::span<char>)
{
    // the caller's buffers are sized for cub; this library's link step has its own temporary layout
    buildOctreeGpu(exec, cstoneTree, d);
}
template void buildOctreeGpu(execution::Gpu, const uint32_t*, OctreeView<uint32_t>, std::span<uint32_t>,
                             std::span<TreeNodeIndex>, std::span<char>);
template void buildOctreeGpu(execution::Gpu, const uint64_t*, OctreeView<uint64_t>, std::span<uint64_t>,
                             std::span<TreeNodeIndex>, std::span<char>);

void upsweepSumGpu(execution::Gpu exec, int numLvl, const TreeNodeIndex* lvlRange, const TreeNodeIndex* childOffsets,
                   LocalIndex* counts)
{
    csCheck(cs_upsweep_sum(numLvl, lvlRange, childOffsets, counts, cudaStream_t(exec)), "upsweepSumGpu");
}

/* ------------------------------------------------------------------------------------------------ halos */

template<class KeyType, class T>
void findHalosGpu(execution::Gpu exec, const KeyType* prefixes, const TreeNodeIndex* childOffsets,
                  const TreeNodeIndex* parents, const Vec3<T>* nodeCenters, const Vec3<T>* nodeSizes,
                  const KeyType* leaves, const Vec3<T>* searchCenters, const Vec3<T>* searchSizes, const Box<T>& box,
                  TreeNodeIndex firstNode, TreeNodeIndex lastNode, uint8_t* collisionFlags)
{
    BoxArgs<T> b(box);
    auto C = [](const Vec3<T>* p) { return reinterpret_cast<const T*>(p); };
    if constexpr (sizeof(KeyType) == 4)
    {
        csCheck(cs_find_halos_u32f(reinterpret_cast<const uint32_t*>(prefixes), childOffsets, parents, C(nodeCenters),
                                   C(nodeSizes), reinterpret_cast<const uint32_t*>(leaves), C(searchCenters),
                                   C(searchSizes), b.lim, b.bnd, firstNode, lastNode, collisionFlags,
                                   cudaStream_t(exec)),
                "findHalosGpu");
    }
    else if constexpr (std::is_same_v<T, float>)
    {
        csCheck(cs_find_halos_u64f(reinterpret_cast<const uint64_t*>(prefixes), childOffsets, parents, C(nodeCenters),
                                   C(nodeSizes), reinterpret_cast<const uint64_t*>(leaves), C(searchCenters),
                                   C(searchSizes), b.lim, b.bnd, firstNode, lastNode, collisionFlags,
                                   cudaStream_t(exec)),
                "findHalosGpu");
    }
    else
    {
        csCheck(cs_find_halos_u64d(reinterpret_cast<const uint64_t*>(prefixes), childOffsets, parents, C(nodeCenters),
                                   C(nodeSizes), reinterpret_cast<const uint64_t*>(leaves), C(searchCenters),
                                   C(searchSizes), b.lim, b.bnd, firstNode, lastNode, collisionFlags,
                                   cudaStream_t(exec)),
                "findHalosGpu");
    }
}
#define CS_FIND_HALOS(KeyType, T)                                                                                      \
    template void findHalosGpu(execution::Gpu, const KeyType*, const TreeNodeIndex*, const TreeNodeIndex*,             \
                               const Vec3<T>*, const Vec3<T>*, const KeyType*, const Vec3<T>*, const Vec3<T>*,          \
                               const Box<T>&, TreeNodeIndex, TreeNodeIndex, uint8_t*);
CS_FIND_HALOS(uint32_t, float)
CS_FIND_HALOS(uint64_t, float)
CS_FIND_HALOS(uint64_t, double)

template<class Tc, class Th>
void computeBoundingBoxGpu(execution::Gpu exec, const Tc* x, const Tc* y, const Tc* z, const Th* h,
                           const LocalIndex* layout, TreeNodeIndex first, TreeNodeIndex last, Th scale,
                           Vec3<Tc>* centers, Vec3<Tc>* sizes)
{
    if constexpr (std::is_same_v<Tc, double> && std::is_same_v<Th, float>)
    {
        csCheck(cs_compute_bounding_boxes_df(x, y, z, h, layout, first, last, scale,
                                             reinterpret_cast<double*>(centers), reinterpret_cast<double*>(sizes),
                                             cudaStream_t(exec)),
                "computeBoundingBoxGpu");
    }
    else if constexpr (std::is_same_v<Tc, float>)
    {
        csCheck(cs_compute_bounding_boxes_f(x, y, z, h, layout, first, last, scale, reinterpret_cast<float*>(centers),
                                            reinterpret_cast<float*>(sizes), cudaStream_t(exec)),
                "computeBoundingBoxGpu");
    }
    else
    {
        csCheck(cs_compute_bounding_boxes_d(x, y, z, h, layout, first, last, scale, reinterpret_cast<double*>(centers),
                                            reinterpret_cast<double*>(sizes), cudaStream_t(exec)),
                "computeBoundingBoxGpu");
    }
}
template void computeBoundingBoxGpu(execution::Gpu, const double*, const double*, const double*, const double*,
                                    const LocalIndex*, TreeNodeIndex, TreeNodeIndex, double, Vec3<double>*,
                                    Vec3<double>*);
template void computeBoundingBoxGpu(execution::Gpu, const double*, const double*, const double*, const float*,
                                    const LocalIndex*, TreeNodeIndex, TreeNodeIndex, float, Vec3<double>*,
                                    Vec3<double>*);
template void computeBoundingBoxGpu(execution::Gpu, const float*, const float*, const float*, const float*,
                                    const LocalIndex*, TreeNodeIndex, TreeNodeIndex, float, Vec3<float>*, Vec3<float>*);

template<class KeyType, class T>
void computeGeoCentersGpu(execution::Gpu exec, const KeyType* prefixes, TreeNodeIndex numNodes, Vec3<T>* centers,
                          Vec3<T>* sizes, const Box<T>& box)
{
    BoxArgs<T> b(box);
    if constexpr (sizeof(KeyType) == 4)
    {
        csCheck(cs_compute_geo_centers_u32f(0, reinterpret_cast<const uint32_t*>(prefixes), numNodes,
                                            reinterpret_cast<float*>(centers), reinterpret_cast<float*>(sizes), b.lim,
                                            b.bnd, cudaStream_t(exec)),
                "computeGeoCentersGpu");
    }
    else if constexpr (std::is_same_v<T, float>)
    {
        csCheck(cs_compute_geo_centers_u64f(0, reinterpret_cast<const uint64_t*>(prefixes), numNodes,
                                            reinterpret_cast<float*>(centers), reinterpret_cast<float*>(sizes), b.lim,
                                            b.bnd, cudaStream_t(exec)),
                "computeGeoCentersGpu");
    }
    else
    {
        csCheck(cs_compute_geo_centers_u64d(0, reinterpret_cast<const uint64_t*>(prefixes), numNodes,
                                            reinterpret_cast<double*>(centers), reinterpret_cast<double*>(sizes), b.lim,
                                            b.bnd, cudaStream_t(exec)),
                "computeGeoCentersGpu");
    }
}
template void computeGeoCentersGpu(execution::Gpu, const uint32_t*, TreeNodeIndex, Vec3<float>*, Vec3<float>*,
                                   const Box<float>&);
template void computeGeoCentersGpu(execution::Gpu, const uint64_t*, TreeNodeIndex, Vec3<float>*, Vec3<float>*,
                                   const Box<float>&);
template void computeGeoCentersGpu(execution::Gpu, const uint64_t*, TreeNodeIndex, Vec3<double>*, Vec3<double>*,
                                   const Box<double>&);

/* ------------------------------------------------------------------------------------------------ groups_gpu.h */

void computeFixedGroups(execution::Gpu exec, LocalIndex first, LocalIndex last, unsigned groupSize,
                        GroupData<execution::Gpu>& groups)
{
    LocalIndex numBodies = last - first;
    LocalIndex numGroups = (numBodies + groupSize - 1) / groupSize;
    groups.data.resize(numGroups + 1);
    csCheck(cs_compute_fixed_groups(first, last, groupSize, rawPtr(groups.data), cudaStream_t(exec)),
            "computeFixedGroups");
    groups.firstBody  = first;
    groups.lastBody   = last;
    groups.numGroups  = numGroups;
    groups.groupStart = rawPtr(groups.data);
    groups.groupEnd   = rawPtr(groups.data) + 1;
}

template<class Tc, class T, class KeyType>
void computeGroupSplits(execution::Gpu exec, LocalIndex first, LocalIndex last, const Tc* x, const Tc* y, const Tc* z,
                        const T* h, const KeyType* leaves, TreeNodeIndex numLeaves, const LocalIndex* layout,
                        const Box<Tc> box, unsigned groupSize, float tolFactor,
                        DeviceVector<LocalIndex>& /*numSplitsPerGroup: scratch of the reference, not needed*/,
                        DeviceVector<LocalIndex>& groups)
{
    static_assert(sizeof(KeyType) == 8, "computeGroupSplits is instantiated for 64-bit keys only (groups_gpu.cu:148-150)");
    if (groupSize != 32 && groupSize != 64) { throw std::runtime_error("Unsupported spatial group size\n"); }
    BoxArgs<Tc> b(box);
    uint32_t numGroups = 0;
    const uint64_t* lv = reinterpret_cast<const uint64_t*>(leaves);
    if constexpr (std::is_same_v<Tc, double> && std::is_same_v<T, double>)
    {
        csCheck(cs_group_splits_begin_dd(first, last, x, y, z, h, lv, numLeaves, layout, b.lim, b.bnd, groupSize,
                                         tolFactor, &numGroups, cudaStream_t(exec)),
                "computeGroupSplits");
    }
    else if constexpr (std::is_same_v<Tc, double>)
    {
        csCheck(cs_group_splits_begin_df(first, last, x, y, z, h, lv, numLeaves, layout, b.lim, b.bnd, groupSize,
                                         tolFactor, &numGroups, cudaStream_t(exec)),
                "computeGroupSplits");
    }
    else
    {
        csCheck(cs_group_splits_begin_ff(first, last, x, y, z, h, lv, numLeaves, layout, b.lim, b.bnd, groupSize,
                                         tolFactor, &numGroups, cudaStream_t(exec)),
                "computeGroupSplits");
    }
    groups.resize(numGroups + 1);
    csCheck(cs_group_splits_finish(first, last, groupSize, rawPtr(groups), cudaStream_t(exec)), "computeGroupSplits");
}

#define CS_GROUP_SPLITS(Tc, T, KeyType)                                                                                \
    template void computeGroupSplits(execution::Gpu, LocalIndex, LocalIndex, const Tc*, const Tc*, const Tc*, const T*,  \
                                     const KeyType*, TreeNodeIndex, const LocalIndex*, const Box<Tc>, unsigned, float,  \
                                     DeviceVector<LocalIndex>&, DeviceVector<LocalIndex>&);
CS_GROUP_SPLITS(double, double, uint64_t)
CS_GROUP_SPLITS(double, float, uint64_t)
CS_GROUP_SPLITS(float, float, uint64_t)

/* ------------------------------------------------------------------------------------------------ collisions_gpu.h
 *                                                                                                  (markMacsGpu) */

template<class T, class KeyType>
void markMacsGpu(execution::Gpu exec, const KeyType* prefixes, const TreeNodeIndex* childOffsets,
                 const TreeNodeIndex* parents, const Vec4<T>* centers, const Box<T>& box, const KeyType* focusNodes,
                 TreeNodeIndex numFocusNodes, bool limitSource, uint8_t* markings)
{
    BoxArgs<T> b(box);
    const T* c4 = reinterpret_cast<const T*>(centers);
    if constexpr (sizeof(KeyType) == 4)
    {
        csCheck(cs_mark_macs_u32f(reinterpret_cast<const uint32_t*>(prefixes), childOffsets, parents, c4, b.lim, b.bnd,
                                  reinterpret_cast<const uint32_t*>(focusNodes), numFocusNodes, int(limitSource),
                                  markings, cudaStream_t(exec)),
                "markMacsGpu");
    }
    else if constexpr (std::is_same_v<T, float>)
    {
        csCheck(cs_mark_macs_u64f(reinterpret_cast<const uint64_t*>(prefixes), childOffsets, parents, c4, b.lim, b.bnd,
                                  reinterpret_cast<const uint64_t*>(focusNodes), numFocusNodes, int(limitSource),
                                  markings, cudaStream_t(exec)),
                "markMacsGpu");
    }
    else
    {
        csCheck(cs_mark_macs_u64d(reinterpret_cast<const uint64_t*>(prefixes), childOffsets, parents, c4, b.lim, b.bnd,
                                  reinterpret_cast<const uint64_t*>(focusNodes), numFocusNodes, int(limitSource),
                                  markings, cudaStream_t(exec)),
                "markMacsGpu");
    }
}

#define CS_MARK_MACS(T, KeyType)                                                                                       \
    template void markMacsGpu(execution::Gpu, const KeyType*, const TreeNodeIndex*, const TreeNodeIndex*,              \
                              const Vec4<T>*, const Box<T>&, const KeyType*, TreeNodeIndex, bool, uint8_t*);
CS_MARK_MACS(float, uint32_t)
CS_MARK_MACS(float, uint64_t)
CS_MARK_MACS(double, uint64_t)

/* ------------------------------------------------------------------------------------------------ gather_halos_gpu.h */

template<class T, class IndexType>
void gatherRanges(execution::Gpu exec, const IndexType* rangeScan, const IndexType* rangeOffsets, int numRanges,
                  const T* src, T* buffer, size_t bufferSize)
{
    static_assert(sizeof(IndexType) == 4 && sizeof(T) % 4 == 0);
    csCheck(cs_gather_ranges(reinterpret_cast<const uint32_t*>(rangeScan), reinterpret_cast<const uint32_t*>(rangeOffsets),
                             numRanges, src, buffer, bufferSize, int(sizeof(T)), cudaStream_t(exec)),
            "gatherRanges");
}

#define CS_GATHER_RANGES(T, IndexType)                                                                                 \
    template void gatherRanges(execution::Gpu, const IndexType*, const IndexType*, int, const T*, T*, size_t);
CS_GATHER_RANGES(int, unsigned)
using ArrF1 = util::array<float, 1>;
using ArrF2 = util::array<float, 2>;
CS_GATHER_RANGES(ArrF1, unsigned)
CS_GATHER_RANGES(ArrF2, unsigned)
CS_GATHER_RANGES(ArrF3, unsigned)
CS_GATHER_RANGES(ArrF4, unsigned)

/* ------------------------------------------------------------------------------------------------ primitives_gpu.h
 *                                                                                                  (minMax) */

template<class T>
std::tuple<T, T> minMax(execution::Gpu exec, const T* first, const T* last)
{
    T mn{}, mx{};
    const size_t n = size_t(last - first);
    if constexpr (std::is_same_v<T, double>) { csCheck(cs_min_max_d(first, n, &mn, &mx, cudaStream_t(exec)), "minMax"); }
    else if constexpr (std::is_same_v<T, float>) { csCheck(cs_min_max_f(first, n, &mn, &mx, cudaStream_t(exec)), "minMax"); }
    else { csCheck(cs_min_max_u32(first, n, &mn, &mx, cudaStream_t(exec)), "minMax"); }
    return std::make_tuple(mn, mx);
}
template std::tuple<double, double> minMax(execution::Gpu, const double*, const double*);
template std::tuple<float, float> minMax(execution::Gpu, const float*, const float*);
template std::tuple<unsigned, unsigned> minMax(execution::Gpu, const unsigned*, const unsigned*);

} // namespace cstone
