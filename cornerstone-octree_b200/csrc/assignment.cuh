/* host-side SFC domain decomposition helpers shared by assignment.cu and domain.cu */
#pragma once

#include <cstdint>
#include <vector>

#include "common.cuh"

namespace csb
{

//! domain/domaindecomp.hpp:57-112: SFC key range, particle count and global-leaf index range of every rank
template<class K>
struct SfcAssignment
{
    std::vector<K> boundaries;        // numRanks + 1 keys
    std::vector<uint32_t> counts;     // particles assigned to each rank
    std::vector<int> treeOffsets;     // numRanks + 1 indices into the global leaf array
    int numRanks() const { return int(counts.size()); }
};

void uniformBins(const uint32_t* counts, size_t numCounts, int numBins, int* bins, uint32_t* binCounts);
template<class K>
SfcAssignment<K> makeSfcAssignment(int numRanks, const std::vector<uint32_t>& counts, const K* leaves);
template<class K>
std::vector<K> initialGlobalTree(int numRanks);
template<class K>
std::vector<K> initialDomainSplits(int numRanks, int level);
template<class K>
std::vector<K> computeSpanningTree(const std::vector<K>& keys);

//! domain/buffer_description.hpp:20-40: [start, end) assigned particles inside a buffer of `size` elements
struct BufferDescription
{
    uint32_t start, end, size;
};

//! buffer_description.hpp:98-108
inline uint32_t exchangeBufferSize(BufferDescription b, uint32_t numPresent, uint32_t numAssigned)
{
    uint32_t numIncoming = numAssigned - numPresent;
    bool fitHead         = b.start >= numIncoming;
    bool fitTail         = b.size - b.end >= numIncoming;
    return (fitHead || fitTail) ? b.size : b.end + numIncoming;
}

//! buffer_description.hpp:110-118
inline uint32_t receiveStart(BufferDescription b, uint32_t numIncoming)
{
    bool fitHead = b.start >= numIncoming;
    return fitHead ? b.start - numIncoming : b.end;
}

//! buffer_description.hpp:120-125
inline void assignedEnvelope(BufferDescription b, uint32_t numIncoming, uint32_t* newStart, uint32_t* newEnd)
{
    bool fitHead = b.start >= numIncoming;
    if (fitHead)
    {
        *newStart = b.start - numIncoming;
        *newEnd   = b.end;
    }
    else
    {
        *newStart = b.start;
        *newEnd   = b.end + numIncoming;
    }
}

} // namespace csb
