/* Host-side SFC domain decomposition: the arithmetic of the reference's GlobalAssignment that runs on the host on every
 * rank and must give identical results everywhere (domain/domaindecomp.hpp:33-127,178-191,
 * domain/buffer_description.hpp:98-125, domain/assignment.hpp:53-74).  O(#global leaves) work, no device code.
 */
#include <algorithm>
#include <numeric>
#include <vector>

#include "assignment.cuh"
#include "cstone_b200.h"
#include "focus.cuh"

namespace csb
{

//! domaindecomp.hpp:33-55; the bin targets use the reference's double arithmetic (hazard H5)
void uniformBins(const uint32_t* counts, size_t numCounts, int numBins, int* bins, uint32_t* binCounts)
{
    std::vector<uint64_t> countScan(numCounts + 1, 0);
    for (size_t i = 0; i < numCounts; ++i)
        countScan[i + 1] = countScan[i] + counts[i];

    auto binCount = double(countScan.back()) / numBins;
    bins[0]       = 0;
    bins[numBins] = int(numCounts);
    for (int i = 1; i < numBins; ++i)
    {
        uint64_t targetCount = uint64_t(i * binCount);
        bins[i] = int(std::lower_bound(countScan.begin(), countScan.end(), targetCount) - countScan.begin());
    }
    for (int i = 1; i < numBins; ++i)
        binCounts[i - 1] = uint32_t(countScan[bins[i]] - countScan[bins[i - 1]]);
    binCounts[numBins - 1] = uint32_t(countScan.back() - countScan[bins[numBins - 1]]);
}

template<class K>
SfcAssignment<K> makeSfcAssignment(int numRanks, const std::vector<uint32_t>& counts, const K* leaves)
{
    SfcAssignment<K> a;
    a.boundaries.resize(numRanks + 1);
    a.counts.resize(numRanks);
    a.treeOffsets.resize(numRanks + 1);
    uniformBins(counts.data(), counts.size(), numRanks, a.treeOffsets.data(), a.counts.data());
    for (int r = 0; r <= numRanks; ++r)
        a.boundaries[r] = leaves[a.treeOffsets[r]];
    return a;
}

//! sfc/common.hpp:119-125
template<class K>
unsigned log8ceil(K n)
{
    if (n == 0) { return 0; }
    unsigned lz = unsigned(clz(K(n - 1)));
    return KeyTraits<K>::maxLevel - (lz - KeyTraits<K>::unusedBits) / 3;
}

//! domaindecomp.hpp:213-228 with enclosingBoxCode (sfc/common.hpp:326-331)
template<class K>
std::vector<K> initialDomainSplits(int numRanks, int level)
{
    std::vector<K> ret(numRanks + 1);
    K delta = nodeRange<K>(0) / K(numRanks);
    ret[0]  = 0;
    for (int i = 1; i < numRanks; ++i)
    {
        K mask = K(nodeRange<K>(unsigned(level)) - 1);
        ret[i] = K((K(i) * delta) & ~mask);
    }
    ret[numRanks] = nodeRange<K>(0);
    return ret;
}

//! tree/csarray.hpp:483-510
template<class K>
std::vector<K> computeSpanningTree(const std::vector<K>& keys)
{
    std::vector<int> offsets(keys.size(), 0);
    for (size_t i = 0; i + 1 < keys.size(); ++i)
        offsets[i + 1] = offsets[i] + spanSfcRangeHost<K>(keys[i], keys[i + 1], nullptr);
    std::vector<K> tree(size_t(offsets.back()) + 1);
    for (size_t i = 0; i + 1 < keys.size(); ++i)
        spanSfcRangeHost<K>(keys[i], keys[i + 1], tree.data() + offsets[i]);
    tree.back() = nodeRange<K>(0);
    return tree;
}

//! assignment.hpp:62-65: the tree every rank starts from
template<class K>
std::vector<K> initialGlobalTree(int numRanks)
{
    unsigned level = log8ceil<K>(K(100) * K(numRanks));
    return computeSpanningTree<K>(initialDomainSplits<K>(numRanks, int(level)));
}

template SfcAssignment<uint32_t> makeSfcAssignment<uint32_t>(int, const std::vector<uint32_t>&, const uint32_t*);
template SfcAssignment<uint64_t> makeSfcAssignment<uint64_t>(int, const std::vector<uint32_t>&, const uint64_t*);
template std::vector<uint32_t> initialGlobalTree<uint32_t>(int);
template std::vector<uint64_t> initialGlobalTree<uint64_t>(int);
template std::vector<uint32_t> initialDomainSplits<uint32_t>(int, int);
template std::vector<uint64_t> initialDomainSplits<uint64_t>(int, int);
template std::vector<uint32_t> computeSpanningTree<uint32_t>(const std::vector<uint32_t>&);
template std::vector<uint64_t> computeSpanningTree<uint64_t>(const std::vector<uint64_t>&);

} // namespace csb

extern "C"
{

int cs_uniform_bins(const uint32_t* counts, size_t numCounts, int numBins, int* bins, uint32_t* binCounts)
{
    CSB_REQUIRE(numBins >= 1, "numBins must be positive");
    csb::uniformBins(counts, numCounts, numBins, bins, binCounts);
    return 0;
}

long cs_initial_global_tree_u32(int numRanks, uint32_t* leaves, long capacity)
{
    auto t = csb::initialGlobalTree<uint32_t>(numRanks);
    if (leaves && long(t.size()) <= capacity) { std::copy(t.begin(), t.end(), leaves); }
    return long(t.size());
}

long cs_initial_global_tree_u64(int numRanks, uint64_t* leaves, long capacity)
{
    auto t = csb::initialGlobalTree<uint64_t>(numRanks);
    if (leaves && long(t.size()) <= capacity) { std::copy(t.begin(), t.end(), leaves); }
    return long(t.size());
}

long cs_spanning_tree_u64(const uint64_t* keys, long numKeys, uint64_t* leaves, long capacity)
{
    auto t = csb::computeSpanningTree<uint64_t>(std::vector<uint64_t>(keys, keys + numKeys));
    if (leaves && long(t.size()) <= capacity) { std::copy(t.begin(), t.end(), leaves); }
    return long(t.size());
}

int cs_exchange_buffer_layout(uint32_t start, uint32_t end, uint32_t size, uint32_t numPresent, uint32_t numAssigned,
                              uint32_t* out4)
{
    csb::BufferDescription b{start, end, size};
    uint32_t numIncoming = numAssigned - numPresent;
    out4[0]              = csb::exchangeBufferSize(b, numPresent, numAssigned);
    csb::BufferDescription e{start, end, out4[0]};
    out4[1] = csb::receiveStart(e, numIncoming);
    csb::assignedEnvelope(e, numIncoming, &out4[2], &out4[3]);
    return 0;
}

} // extern "C"
