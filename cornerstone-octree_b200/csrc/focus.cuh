/* declarations shared between focus.cu, csarray.cu, octree.cu, ... and the Domain driver (domain.cu) */
#pragma once

#include "common.cuh"

namespace csb
{

enum EnforceStatus : int
{
    ENFORCE_CONVERGED    = 0,
    ENFORCE_CANCEL_MERGE = 1,
    ENFORCE_REBALANCE    = 2,
    ENFORCE_FAILED       = 3
};

/* focus.cu */
template<class K>
int essentialOps(const K* prefixes, const int* childOffsets, const int* parents, const uint32_t* counts,
                 const uint8_t* macs, K focusStart, K focusEnd, uint32_t bucketSize, int* nodeOps, int numNodes,
                 cudaStream_t s);
template<class K>
int enforceKeys(const K* keys, int numKeys, const K* prefixes, const int* childOffsets, const int* parents,
                int* nodeOps, int* statusDev, cudaStream_t s);
template<class K>
int protectAncestors(const K* prefixes, const int* parents, int* nodeOps, int numNodes, int* changesDev,
                     cudaStream_t s);
int gatherLeafOps(const int* leafToInternalLeaves, int numLeaves, const int* nodeOpsAll, int* leafOps,
                  int* notAllOneDev, cudaStream_t s);
template<class K>
int countGaps(const K* keys, int numGaps, uint32_t* gapCounts, cudaStream_t s);
template<class K>
int fillGaps(const K* keys, int numGaps, const uint32_t* offsets, K* out, cudaStream_t s);
int scatterCounts(const int* leafToInternalLeaves, int numLeaves, const uint32_t* leafCounts, uint32_t* nodeCounts,
                  cudaStream_t s);
template<class T>
int gatherVec3(const int* map, int n, const T* src, T* dst, cudaStream_t s);
int layoutCounts(const uint32_t* leafCounts, const uint8_t* flags, const int* leafToInternalLeaves, int numLeaves,
                 int ownStart, int ownEnd, uint32_t* layout, cudaStream_t s);
template<class T>
int minMaxPartials(const T* a, size_t n, T* partial, int numBlocks, cudaStream_t s);
int maxU32(const uint32_t* a, size_t n, uint32_t* resultDev, cudaStream_t s);
//! a[i] = max(a[i], b[i])
int maxInto(uint32_t* a, const uint32_t* b, size_t n, cudaStream_t s);
template<class K>
int lowerBounds(const K* keys, size_t n, const K* targets, int numTargets, uint32_t* out, cudaStream_t s);
template<class K>
int spanSfcRangeHost(K a, K b, K* output);

/* let.cu */
template<class T>
int minMacCenters(const T* geoCenters, const T* geoSizes, int numNodes, float invThetaEff, T* centers4, cudaStream_t s);
template<class K, class T>
int markMacs(const K* prefixes, const int* childOffsets, const int* parents, const T* centers4, const double* lim,
             const int* bnd, const K* focusNodes, int numFocusNodes, uint8_t* markings, cudaStream_t s, bool limitSource = false);
template<class K>
int rangeCount(const K* gLeaves, int numGlobalLeaves, const uint64_t* gCountScan, const K* fLeaves, const int* idx,
               int numIdx, uint32_t* leafCounts, cudaStream_t s);
template<class K>
int focusBounds(const K* leaves, int numKeys, const K* bounds, int numBounds, int* out, cudaStream_t s);
template<class K>
int notIncluded(const K* fLeaves, int count, const K* gLeaves, int gCount, int* flag, cudaStream_t s);
template<class K>
int checkTreelet(const K* treelet, int count, const K* leaves, int numLeaves, uint32_t* valid, cudaStream_t s);
template<class K>
int splitTreelet(const K* keys, int count, const uint32_t* validScan, K* accepted, K* rejected, cudaStream_t s);
template<class K>
int rejectLeaves(const K* rejected, int count, const K* leaves, int numKeys, int* nodeOps, cudaStream_t s);
int fillInt(int* a, int n, int v, cudaStream_t s);
template<class K>
int indexTreelet(const K* treelet, int numNodes, const K* prefixes, const int* levelRange, int* out, int* error,
                 cudaStream_t s);
int haloRunStarts(const uint32_t* layout, int numLeaves, const int* fa, int numRanks, int me, uint32_t maxParticles,
                  uint32_t* flags, int* status, cudaStream_t s);
template<class K>
int haloRequestKeys(const uint32_t* layout, int numLeaves, const int* fa, int numRanks, int me,
                    const uint32_t* startScan, const K* leaves, K* req, cudaStream_t s);
template<class K>
int haloRanges(const K* pairs, int numPairs, const K* leaves, int numKeys, const uint32_t* layout, uint32_t* len,
               uint32_t* start, cudaStream_t s);
int pickU32(const uint32_t* src, const int* idx, int n, uint32_t* dst, cudaStream_t s);
int gatherU32(const int* idx, int n, const uint32_t* src, uint32_t* dst, cudaStream_t s);
int scatterU32(const int* idx, int n, const uint32_t* src, uint32_t* dst, cudaStream_t s);
int gatherRangesWords(const uint32_t* rangeScan, const uint32_t* rangeStart, int numRanges, uint32_t total, int words,
                      const void* src, void* out, cudaStream_t s);
//! (x,y,z,h) records of the elements [first, first + n) at the same positions of `rec` (sort.cu)
int packRecords4(const void* const* src4, size_t first, size_t n, void* rec, int elemBytes, cudaStream_t s);
//! dst4[k][i] = rec[ordering[i]].v[k]; the destination arrays may be peer memory (sort.cu)
int gatherFromRecords4(const uint32_t* ordering, size_t n, const void* rec, void* const* dst4, int elemBytes,
                       cudaStream_t s);
//! out[k] = src[order[k]], elements of `words` 32-bit words
int gatherWords(const uint32_t* order, uint32_t n, int words, const void* src, void* out, cudaStream_t s);
//! replay of a recorded particle exchange for one more field (reapplySync, domain/domain.hpp:297-329)
int replayGatherWords(const uint32_t* order, uint32_t n, int words, const void* before, const void* received,
                      uint32_t recvStart, uint32_t numRecv, void* out, cudaStream_t s);
template<class E>
int gatherRanges4(const uint32_t* rangeScan, const uint32_t* rangeStart, int numRanges, uint32_t total, const E* a,
                  const E* b, const E* c, const E* d, E* out, size_t blockElems, cudaStream_t s);

/* csarray.cu */
template<class K>
int computeNodeCounts(const K* leaves, uint32_t* counts, int numLeaves, const K* keys, size_t n, uint32_t maxCount,
                      cudaStream_t s);
template<class K>
int computeNodeOps(const K* leaves, int numLeaves, const uint32_t* counts, uint32_t bucketSize, int* nodeOps, void* tmp,
                   int* newNumLeaves, int* converged, cudaStream_t s);
template<class K>
int rebalanceTree(const K* leaves, int numLeaves, int newNumLeaves, const int* nodeOps, K* newLeaves, cudaStream_t s);
size_t nodeOpsTempBytes(size_t numLeaves);

/* octree.cu */
template<class K>
int buildOctree(const K* leaves, int numLeaves, K* prefixes, int* childOffsets, int* parents, int* levelRange,
                int* internalToLeaf, int* leafToInternal, void* tmp, size_t tmpBytes, cudaStream_t s);
template<class K, class T>
int computeGeoCenters(int kind, const K* prefixes, int numNodes, T* centers, T* sizes, const double* lim,
                      const int* bnd, cudaStream_t s);
int upsweepSum(int maxLevel, const int* levelRangeHost, const int* childOffsets, uint32_t* counts, cudaStream_t s);
size_t buildOctreeTempBytesU32(int numLeaves);
size_t buildOctreeTempBytesU64(int numLeaves);

/* merge.cu */
template<class K>
int mergeSortedRuns(K* keys, uint32_t* vals, const size_t* runOffsets, int numRuns, K* keyBuf, uint32_t* valBuf,
                    cudaStream_t s, bool* resultInBuffers = nullptr);

/* sort.cu */
int sortByKeyU64(uint64_t*, uint32_t*, size_t, uint64_t*, uint32_t*, void*, size_t, cudaStream_t);
int sortByKeyU32(uint32_t*, uint32_t*, size_t, uint32_t*, uint32_t*, void*, size_t, cudaStream_t);
int sortByKeyIotaU64(uint64_t*, uint32_t*, uint32_t, size_t, uint64_t*, uint32_t*, void*, size_t, cudaStream_t);
int sortByKeyIotaU32(uint32_t*, uint32_t*, uint32_t, size_t, uint32_t*, uint32_t*, void*, size_t, cudaStream_t);
size_t sortTempBytesU64(size_t n);
size_t sortTempBytesU32(size_t n);

/* halos.cu / neighbors.cu */
template<class T, class Th>
int computeBoundingBoxes(const T* x, const T* y, const T* z, const Th* h, const uint32_t* layout, int firstLeaf,
                         int lastLeaf, Th scale, T* sc, T* ss, cudaStream_t s);
template<class K, class T>
int findHalos(const K* prefixes, const int* childOffsets, const int* parents, const T* centers, const T* sizes,
              const K* leaves, const T* searchCenters, const T* searchSizes, const double* lim, const int* bnd,
              int firstLeaf, int lastLeaf, uint8_t* flags, cudaStream_t s);
template<class T, class Th>
int findNeighbors(const T* x, const T* y, const T* z, const Th* h, uint32_t first, uint32_t last, const double* lim,
                  const int* bnd, int numLeaves, const int* childOffsets, const int* parents,
                  const int* internalToLeaf, const uint32_t* layout, const T* centers, const T* sizes, uint32_t ngmax,
                  uint32_t* neighbors,
                  uint32_t* neighborsCount, cudaStream_t s, float searchExtFactor = 1.0f);

} // namespace csb
