/* cornerstone-b200 C ABI — the drop-in boundary for the Cornerstone domain-sync hot path on B200 (sm_100a).
 *
 * Every entry point replaces one GPU entry point of the reference (sekelle/cornerstone-octree); the reference
 * declaration it stands in for is cited as file:line relative to /root/reference/include/cstone.  INTEGRATION.md
 * shows the forwarding stubs a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain C: pointers + sizes only. Unless stated otherwise all array pointers are DEVICE pointers, `stream` is a
 *    cudaStream_t passed as void* (NULL = default stream) and work is enqueued asynchronously on it.
 *  - functions whose results are needed on the host (sizes, convergence flags) synchronise the stream internally and
 *    say so.
 *  - suffixes: key type u32|u64, coordinate type f|d: *_u32f (uint32_t,float) *_u64f (uint64_t,float)
 *    *_u64d (uint64_t,double).  `kind`: 0 = Hilbert, 1 = Morton.
 *  - box: `lim` = host pointer to {xmin,xmax,ymin,ymax,zmin,zmax} (double, converted to the coordinate type exactly
 *    like cstone::Box<T>'s constructor, sfc/box.hpp:100-122), `bnd` = host pointer to 3 ints, BoundaryType values
 *    0 open, 1 periodic, 2 fixed, 3 cubic_open (sfc/box.hpp:78-84).
 *  - return value: 0 on success; non-zero on failure with a message retrievable through cs_last_error().  The C++
 *    forwarders turn CUDA failures into the reference's print+exit behaviour (cuda/errorcheck.cuh:15-27) and
 *    contract violations into std::runtime_error (primitives_gpu.cu:338).
 *  - no CPU fallback exists: without a CUDA device every compute call fails with a non-zero status.
 */
#ifndef CSTONE_B200_H
#define CSTONE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

/* ---- library ---- */
const char* cs_last_error(void);
int cs_version(void);
/* number of CUDA kernels launched by this library in the calling process since load (for bench.py's gpu_launches) */
uint64_t cs_kernel_launch_count(void);
/* Entry points that need internal temporaries (the reference takes them from cudaMallocAsync, e.g.
 * primitives_gpu.cu:289,302) use library-owned scratch that is keyed per (device, stream) and only grows;
 * cs_release_scratch() synchronises the devices involved and frees all of it. */
int cs_release_scratch(void);

/* ---- SFC keys: computeSfcKeys(Gpu, x,y,z, keys, n, box)  sfc/sfc_gpu.h:24-26, sfc/sfc_gpu.cu:23-62 ----
 * keys[i] = sfc3D(x[i],y[i],z[i], box) unless keys[i] == removeKey = 2^(3*maxTreeLevel) (sfc/sfc.hpp:274). */
int cs_compute_sfc_keys_u32f(int kind, const float* x, const float* y, const float* z, uint32_t* keys, size_t n,
                             const double* lim, const int* bnd, void* stream);
int cs_compute_sfc_keys_u32d(int kind, const double* x, const double* y, const double* z, uint32_t* keys, size_t n,
                             const double* lim, const int* bnd, void* stream);
int cs_compute_sfc_keys_u64f(int kind, const float* x, const float* y, const float* z, uint64_t* keys, size_t n,
                             const double* lim, const int* bnd, void* stream);
int cs_compute_sfc_keys_u64d(int kind, const double* x, const double* y, const double* z, uint64_t* keys, size_t n,
                             const double* lim, const int* bnd, void* stream);

/* ---- key+index sort: sortByKey(Gpu, first,last, values, keyBuf, valueBuf, tmp, tmpBytes)
 *      primitives/primitives_gpu.h:115-155, primitives_gpu.cu:310-356 ----
 * Stable ascending LSD radix sort of (key, uint32 value) pairs over all key bits; the result is left in
 * keys/values (values may be NULL for a keys-only sort).  keyBuf/valueBuf: n elements each; tmp: at least
 * cs_sort_by_key_temp_bytes_*(n) bytes.  n < 2^30. */
size_t cs_sort_by_key_temp_bytes_u32(size_t n);
size_t cs_sort_by_key_temp_bytes_u64(size_t n);
int cs_sort_by_key_u32(uint32_t* keys, uint32_t* values, size_t n, uint32_t* keyBuf, uint32_t* valueBuf, void* tmp,
                       size_t tmpBytes, void* stream);
int cs_sort_by_key_u64(uint64_t* keys, uint32_t* values, size_t n, uint64_t* keyBuf, uint32_t* valueBuf, void* tmp,
                       size_t tmpBytes, void* stream);

/* sequence(Gpu, start, n, out)  primitives_gpu.h:60-63 : out[i] = start + i */
int cs_sequence_u32(uint32_t start, size_t n, uint32_t* out, void* stream);

/* gather(Gpu, ordering, src, dst): dst[i] = src[ordering[i]]  primitives_gpu.h:30-42, primitives_gpu.cu:74-90.
 * elemBytes in {4, 8}. */
int cs_gather(const uint32_t* ordering, size_t n, const void* src, void* dst, int elemBytes, void* stream);
/* fused form of gatherArrays(x,y,z,h) (domain/layout.hpp:230-261): one ordering read feeds four payload streams */
int cs_gather4(const uint32_t* ordering, size_t n, const void* const* src4, void* const* dst4, int elemBytes,
               void* stream);

/* gatherArrays(gatherFunc, ordering, n, inOffset = 0, outOffset = 0, {x,y,z,h}, scratch) (domain/layout.hpp:230-261)
 * for four arrays of srcCount elements each (every ordering[i] < srcCount): dst4[k][i] = src4[k][ordering[i]].  Uses
 * library scratch (4 * elemBytes * srcCount) to turn four scattered reads per particle into one record read. */
int cs_gather_arrays4(const uint32_t* ordering, size_t n, size_t srcCount, const void* const* src4, void* const* dst4,
                      int elemBytes, void* stream);

/* exclusiveScan(Gpu, in, in+n, out)  primitives_gpu.h:66-72 ; tmp >= cs_scan_temp_bytes(n) */
size_t cs_scan_temp_bytes(size_t n);
int cs_exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void* tmp, void* stream);

/* ---- cornerstone leaf array (tree/csarray_gpu.h:41-82, csarray_gpu.cu) ----
 * computeNodeCountsGpu: counts[i] = min(#keys in [leaves[i], leaves[i+1]), maxCount); keys sorted ascending */
int cs_compute_node_counts_u32(const uint32_t* leaves, uint32_t* counts, int numLeaves, const uint32_t* keys, size_t n,
                               uint32_t maxCount, void* stream);
int cs_compute_node_counts_u64(const uint64_t* leaves, uint32_t* counts, int numLeaves, const uint64_t* keys, size_t n,
                               uint32_t maxCount, void* stream);
/* computeNodeOpsGpu (csarray_gpu.cu:182-205): nodeOps[numLeaves+1] <- exclusive scan of the rebalance decisions
 * (0 merge, 1 keep, 8/64/512/4096 split).  Synchronises; *newNumLeaves (host) = nodeOps[numLeaves],
 * *converged (host) = 1 iff every decision was 1.  tmp >= cs_node_ops_temp_bytes(numLeaves). */
size_t cs_node_ops_temp_bytes(int numLeaves);
int cs_compute_node_ops_u32(const uint32_t* leaves, int numLeaves, const uint32_t* counts, uint32_t bucketSize,
                            int* nodeOps, void* tmp, int* newNumLeaves, int* converged, void* stream);
int cs_compute_node_ops_u64(const uint64_t* leaves, int numLeaves, const uint32_t* counts, uint32_t bucketSize,
                            int* nodeOps, void* tmp, int* newNumLeaves, int* converged, void* stream);
/* rebalanceTreeGpu (csarray_gpu.cu:212-231): newLeaves[newNumLeaves+1] from scanned nodeOps */
int cs_rebalance_tree_u32(const uint32_t* leaves, int numLeaves, int newNumLeaves, const int* nodeOps,
                          uint32_t* newLeaves, void* stream);
int cs_rebalance_tree_u64(const uint64_t* leaves, int numLeaves, int newNumLeaves, const int* nodeOps,
                          uint64_t* newLeaves, void* stream);
/* countSfcGapsGpu / fillSfcGapsGpu (tree/csarray_gpu.h:78-82, csarray_gpu.cu:238-270): nodeOps[i] = number of octree
 * nodes that span [tree[i], tree[i+1]) (nodeOps has numNodes + 1 entries, the last is set to 0); after an exclusive
 * scan, fill writes those nodes to newTree + nodeOps[i] and tree[numNodes] to newTree[nodeOps[numNodes]] */
int cs_count_sfc_gaps_u32(const uint32_t* tree, int numNodes, int* nodeOps, void* stream);
int cs_count_sfc_gaps_u64(const uint64_t* tree, int numNodes, int* nodeOps, void* stream);
int cs_fill_sfc_gaps_u32(const uint32_t* tree, int numNodes, const int* nodeOps, uint32_t* newTree, void* stream);
int cs_fill_sfc_gaps_u64(const uint64_t* tree, int numNodes, const int* nodeOps, uint64_t* newTree, void* stream);
/* computeOctree (csarray.hpp:429-440 / test/unit_cuda/tree/csarray.cu:164-185): converge a leaf array from the root
 * for sorted keys.  leaves: capacity+1 keys, counts: capacity.  Synchronises; returns the leaf count in
 * *numLeaves (host), status 3 if capacity is too small (then *numLeaves = required). */
int cs_compute_octree_u32(const uint32_t* keys, size_t n, uint32_t bucketSize, uint32_t* leaves, uint32_t* counts,
                          int capacity, int* numLeaves, void* stream);
int cs_compute_octree_u64(const uint64_t* keys, size_t n, uint32_t bucketSize, uint64_t* leaves, uint32_t* counts,
                          int capacity, int* numLeaves, void* stream);

/* ---- internal octree: buildOctreeGpu(cstoneTree, OctreeView)  tree/octree_gpu.h:35-52, octree_gpu.cu:139-168 ----
 * OctreeData layout (tree/octree.hpp:285-360): prefixes[numNodes], childOffsets[numNodes] (0 = leaf),
 * parents[(numNodes-1)/8], levelRange[maxTreeLevel+2] (device), internalToLeaf[numNodes], leafToInternal[numNodes]
 * with numInternal = (numLeaves-1)/7, numNodes = numLeaves+numInternal.  tmp >= cs_build_octree_temp_bytes_*. */
size_t cs_build_octree_temp_bytes_u32(int numLeaves);
size_t cs_build_octree_temp_bytes_u64(int numLeaves);
int cs_build_octree_u32(const uint32_t* leaves, int numLeaves, uint32_t* prefixes, int* childOffsets, int* parents,
                        int* levelRange, int* internalToLeaf, int* leafToInternal, void* tmp, size_t tmpBytes,
                        void* stream);
int cs_build_octree_u64(const uint64_t* leaves, int numLeaves, uint64_t* prefixes, int* childOffsets, int* parents,
                        int* levelRange, int* internalToLeaf, int* leafToInternal, void* tmp, size_t tmpBytes,
                        void* stream);
/* upsweepSumGpu (octree_gpu.h:66-71, octree_gpu.cu:205-236): counts[node] = min(sum of 8 children, 2^32-1);
 * levelRange is a HOST pointer here (as in the reference) */
int cs_upsweep_sum(int maxLevel, const int* levelRangeHost, const int* childOffsets, uint32_t* counts, void* stream);

/* computeGeoCentersGpu (focus/source_center_gpu.h:100-106, source_center_gpu.cu:212-237): centers/sizes are
 * Vec3<T>[numNodes] (3 consecutive T).  kind selects the curve used to decode node boxes (the reference always uses
 * Hilbert through SfcKind, sfc/sfc.hpp:39-40). */
int cs_compute_geo_centers_u32f(int kind, const uint32_t* prefixes, int numNodes, float* centers, float* sizes,
                                const double* lim, const int* bnd, void* stream);
int cs_compute_geo_centers_u64f(int kind, const uint64_t* prefixes, int numNodes, float* centers, float* sizes,
                                const double* lim, const int* bnd, void* stream);
int cs_compute_geo_centers_u64d(int kind, const uint64_t* prefixes, int numNodes, double* centers, double* sizes,
                                const double* lim, const int* bnd, void* stream);

/* ---- halo discovery ----
 * computeBoundingBoxGpu (source_center_gpu.h:40-52, source_center_gpu.cu:23-92): per leaf in [firstLeaf,lastLeaf)
 * the AABB of particle spheres of radius h*scale, initialised with searchCenters[leaf] */
int cs_compute_bounding_boxes_f(const float* x, const float* y, const float* z, const float* h, const uint32_t* layout,
                                int firstLeaf, int lastLeaf, float scale, float* searchCenters, float* searchSizes,
                                void* stream);
int cs_compute_bounding_boxes_d(const double* x, const double* y, const double* z, const double* h,
                                const uint32_t* layout, int firstLeaf, int lastLeaf, double scale,
                                double* searchCenters, double* searchSizes, void* stream);
/* mixed precision: double coordinates, float smoothing lengths (the (double, float) instantiation of
 * source_center_gpu.cu:91): r = h * scale is formed in float and promoted, as in the reference */
int cs_compute_bounding_boxes_df(const double* x, const double* y, const double* z, const float* h,
                                 const uint32_t* layout, int firstLeaf, int lastLeaf, float scale,
                                 double* searchCenters, double* searchSizes, void* stream);
/* findHalosGpu (traversal/collisions_gpu.h:46-58, collisions_gpu.cu:23-88): flags[numNodes] must be zeroed by the
 * caller unless accumulating */
int cs_find_halos_u32f(const uint32_t* prefixes, const int* childOffsets, const int* parents, const float* centers,
                       const float* sizes, const uint32_t* leaves, const float* searchCenters,
                       const float* searchSizes, const double* lim, const int* bnd, int firstLeaf, int lastLeaf,
                       uint8_t* flags, void* stream);
int cs_find_halos_u64f(const uint64_t* prefixes, const int* childOffsets, const int* parents, const float* centers,
                       const float* sizes, const uint64_t* leaves, const float* searchCenters,
                       const float* searchSizes, const double* lim, const int* bnd, int firstLeaf, int lastLeaf,
                       uint8_t* flags, void* stream);
int cs_find_halos_u64d(const uint64_t* prefixes, const int* childOffsets, const int* parents, const double* centers,
                       const double* sizes, const uint64_t* leaves, const double* searchCenters,
                       const double* searchSizes, const double* lim, const int* bnd, int firstLeaf, int lastLeaf,
                       uint8_t* flags, void* stream);

/* ---- neighbour search: findNeighbors(x,y,z,h, firstId,lastId, box, OctreeNsView, ngmax, neighbors, neighborsCount)
 *      findneighbors.hpp:156-177 (the reference has no GPU function producing this layout) ----
 * neighbors[(i-firstId)*ngmax + k] (ascending particle index, truncated at ngmax), neighborsCount[i-firstId]
 * (not truncated).  Tree arrays as in OctreeNsView (tree/octree.hpp:259-283); layout[numLeaves+1]. */
int cs_find_neighbors_f(const float* x, const float* y, const float* z, const float* h, uint32_t firstId,
                        uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                        const int* parents, const int* internalToLeaf, const uint32_t* layout, const float* centers,
                        const float* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                        void* stream);
int cs_find_neighbors_d(const double* x, const double* y, const double* z, const double* h, uint32_t firstId,
                        uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                        const int* parents, const int* internalToLeaf, const uint32_t* layout, const double* centers,
                        const double* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                        void* stream);
/* double coordinates, float smoothing lengths (Th != Tc, findneighbors.hpp:89-99) */
int cs_find_neighbors_df(const double* x, const double* y, const double* z, const float* h, uint32_t firstId,
                         uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                         const int* parents, const int* internalToLeaf, const uint32_t* layout, const double* centers,
                         const double* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                         void* stream);
/* the same with OctreeNsView::searchExtFactor (tree/octree.hpp:279-282): the continuation tests of the walk use the
 * radius 2h * searchExtFactor (findneighbors.hpp:100), acceptance stays at 2h */
int cs_find_neighbors_ext_f(const float* x, const float* y, const float* z, const float* h, uint32_t firstId,
                            uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                            const int* parents, const int* internalToLeaf, const uint32_t* layout, const float* centers,
                            const float* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                            float searchExtFactor, void* stream);
int cs_find_neighbors_ext_d(const double* x, const double* y, const double* z, const double* h, uint32_t firstId,
                            uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                            const int* parents, const int* internalToLeaf, const uint32_t* layout,
                            const double* centers, const double* sizes, uint32_t ngmax, uint32_t* neighbors,
                            uint32_t* neighborsCount, float searchExtFactor, void* stream);
int cs_find_neighbors_ext_df(const double* x, const double* y, const double* z, const float* h, uint32_t firstId,
                             uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                             const int* parents, const int* internalToLeaf, const uint32_t* layout,
                             const double* centers, const double* sizes, uint32_t ngmax, uint32_t* neighbors,
                             uint32_t* neighborsCount, float searchExtFactor, void* stream);

/* ---- focus-tree (LET) rebalance decisions: focus/rebalance_gpu.h:27-79 ----
 * rebalanceDecisionEssentialGpu: nodeOps[numNodes] from node counts and MAC flags of the fully linked tree, for the focus
 *   [focusStart, focusEnd); protectAncestorsGpu: nodes whose ancestors change are reset, *converged as the reference
 *   returns it; enforceKeysGpu: makes the mandatory keys resolvable, *status = ResolutionStatus (0 converged,
 *   1 cancelMerge, 2 rebalance, 3 failed; focus/rebalance.hpp:171-178) */
int cs_rebalance_decision_essential_u32(const uint32_t* prefixes, const int* childOffsets, const int* parents,
                                        const uint32_t* counts, const uint8_t* macs, uint32_t focusStart,
                                        uint32_t focusEnd, uint32_t bucketSize, int* nodeOps, int numNodes, void* stream);
int cs_rebalance_decision_essential_u64(const uint64_t* prefixes, const int* childOffsets, const int* parents,
                                        const uint32_t* counts, const uint8_t* macs, uint64_t focusStart,
                                        uint64_t focusEnd, uint32_t bucketSize, int* nodeOps, int numNodes, void* stream);
int cs_protect_ancestors_u32(const uint32_t* prefixes, const int* parents, int* nodeOps, int numNodes, int* converged,
                             void* stream);
int cs_protect_ancestors_u64(const uint64_t* prefixes, const int* parents, int* nodeOps, int numNodes, int* converged,
                             void* stream);
int cs_enforce_keys_u32(const uint32_t* keys, int numKeys, const uint32_t* prefixes, const int* childOffsets,
                        const int* parents, int* nodeOps, int* status, void* stream);
int cs_enforce_keys_u64(const uint64_t* keys, int numKeys, const uint64_t* prefixes, const int* childOffsets,
                        const int* parents, int* nodeOps, int* status, void* stream);

/* ---- extractMarkedElements (domain/layout.hpp:110-141): request keys of the leaves [firstReqIdx, secondReqIdx) that hold
 *      halo particles - one (first key, end key) pair per run of consecutive leaves with layout[i+1] > layout[i].
 *      Returns the number of keys written to `out` (device), -needed if capacity is too small, -1 on error. */
long cs_extract_marked_elements_u32(const uint32_t* leaves, const uint32_t* layout, int numLeaves, int firstReqIdx,
                                    int secondReqIdx, uint32_t* out, long capacity, void* stream);
long cs_extract_marked_elements_u64(const uint64_t* leaves, const uint32_t* layout, int numLeaves, int firstReqIdx,
                                    int secondReqIdx, uint64_t* out, long capacity, void* stream);

/* ---- stable merge of sorted runs: what the second sortByKey of GlobalAssignment::distribute (domain/assignment.hpp:197-201)
 *      amounts to when the present particles and the block of every source rank are already sorted.  runOffsets is a
 *      HOST array of numRuns + 1 element offsets; keyBuf / valueBuf are double buffers of runOffsets[numRuns] elements.
 *      Equal keys keep run order, then their order inside the run (= std::stable_sort of the concatenation). */
int cs_merge_sorted_runs_u32(uint32_t* keys, uint32_t* values, const size_t* runOffsets, int numRuns, uint32_t* keyBuf,
                             uint32_t* valueBuf, void* stream);
int cs_merge_sorted_runs_u64(uint64_t* keys, uint32_t* values, const size_t* runOffsets, int numRuns, uint64_t* keyBuf,
                             uint32_t* valueBuf, void* stream);

/* ---- host-side SFC domain decomposition (no device work; identical on every rank) ----
 * uniformBins (domain/domaindecomp.hpp:33-55): bins[numBins+1] leaf indices with ~equal particle sums, binCounts[numBins] */
int cs_uniform_bins(const uint32_t* counts, size_t numCounts, int numBins, int* bins, uint32_t* binCounts);
/* the global tree every rank starts from: computeSpanningTree(initialDomainSplits(numRanks, log8ceil(100 numRanks)))
 * (domain/assignment.hpp:62-65); returns the number of keys, writes them if capacity suffices */
long cs_initial_global_tree_u32(int numRanks, uint32_t* leaves, long capacity);
long cs_initial_global_tree_u64(int numRanks, uint64_t* leaves, long capacity);
/* computeSpanningTree (tree/csarray.hpp:483-510) */
long cs_spanning_tree_u64(const uint64_t* keys, long numKeys, uint64_t* leaves, long capacity);
/* domain_exchange::{exchangeBufferSize, receiveStart, assignedEnvelope} (domain/buffer_description.hpp:98-125):
 * out4 = {exchangeSize, receiveStart, envelopeStart, envelopeEnd} */
int cs_exchange_buffer_layout(uint32_t start, uint32_t end, uint32_t size, uint32_t numPresent, uint32_t numAssigned,
                              uint32_t* out4);

/* ---- communicators: what MPI_Comm is to the reference's Domain (domain.hpp:63-86) ----
 * local: ranks are threads of one process (any rank -> device mapping); nccl: one process per GPU, libnccl is
 * dlopen()ed on first use.  cs_nccl_unique_id fills 128 bytes that rank 0 distributes to the other ranks. */
typedef struct cs_comm cs_comm_t;
void* cs_local_world_create(int size);
void cs_local_world_destroy(void* world);
/* a rank that fails calls this so that the ranks waiting for it return an error instead of blocking forever */
void cs_local_world_abort(void* world);
cs_comm_t* cs_comm_create_local(void* world, int rank);
int cs_nccl_unique_id(void* out128);
cs_comm_t* cs_comm_create_nccl(int rank, int size, const void* id128);
void cs_comm_destroy(cs_comm_t* comm);
int cs_comm_rank(const cs_comm_t* comm);
int cs_comm_size(const cs_comm_t* comm);
uint64_t cs_comm_bytes_sent(const cs_comm_t* comm);

/* ---- Domain: cstone::Domain<KeyType,T,Gpu> (domain/domain.hpp:38-664) ----
 * cs_domain_create_*  <-> Domain(exec, rank, nRanks, bucketSize, bucketSizeFocus, theta, comm, box)  domain.hpp:63-86
 *                         (returns NULL + cs_last_error() where the reference throws std::runtime_error)
 * cs_domain_sync      <-> sync(keys, x, y, z, h, {}, scratch)                                       domain.hpp:169-218
 * cs_domain_info      <-> startIndex/endIndex/nParticlesWithHalos/box()                             domain.hpp:339-361
 * cs_domain_ptr       <-> the arrays sync() leaves behind + focusTree()/globalTree()/layout()/octreeProperties()
 * cs_domain_find_neighbors <-> findNeighbors(x,y,z,h, startIndex, endIndex, box, octreeProperties(), ...)
 *
 * The domain owns the particle arrays in HBM (a C ABI cannot resize the caller's vectors).  cs_domain_sync takes
 * x,y,z,h (and optionally keys, for removeKey marking) of n particles from device (hostInput = 0) or host
 * (hostInput = 1) memory; passing x == NULL re-synchronises the domain-owned arrays in place after the caller has
 * updated them through cs_domain_ptr.  After the call the arrays hold nParticlesWithHalos elements, SFC-sorted
 * assigned particles in [startIndex, endIndex).  Synchronises the stream.
 * Multi-rank domains (numRanks > 1) need cs_domain_attach_comm before the first sync. */
typedef struct cs_domain cs_domain_t;

enum cs_domain_field
{
    CS_FIELD_X = 0,
    CS_FIELD_Y,
    CS_FIELD_Z,
    CS_FIELD_H,
    CS_FIELD_KEYS,
    CS_FIELD_FOCUS_LEAVES,      /* KeyType[numFocusLeaves + 1]      focusTree().treeLeavesAcc() */
    CS_FIELD_FOCUS_LEAF_COUNTS, /* unsigned[numFocusLeaves]         focusTree().leafCountsAcc() */
    CS_FIELD_FOCUS_NODE_COUNTS, /* unsigned[numFocusNodes]          focusTree().countsAcc()     */
    CS_FIELD_LAYOUT,            /* LocalIndex[numFocusLeaves + 1]   layout()                    */
    CS_FIELD_PREFIXES,          /* OctreeView of the focus tree: tree/octree.hpp:234-256 */
    CS_FIELD_CHILD_OFFSETS,
    CS_FIELD_PARENTS,
    CS_FIELD_LEVEL_RANGE,
    CS_FIELD_INTERNAL_TO_LEAF,
    CS_FIELD_LEAF_TO_INTERNAL,
    CS_FIELD_GEO_CENTERS, /* Vec3<T>[numFocusNodes] */
    CS_FIELD_GEO_SIZES,
    CS_FIELD_HALO_FLAGS,    /* uint8_t[numFocusNodes]            focusTree().flags() */
    CS_FIELD_GLOBAL_LEAVES, /* KeyType[numGlobalLeaves + 1]      globalTree().leaves */
    CS_FIELD_GLOBAL_COUNTS,
    CS_FIELD_GLOBAL_PREFIXES,
    CS_FIELD_GLOBAL_CHILD_OFFSETS
};

cs_domain_t* cs_domain_create_u32f(int rank, int numRanks, unsigned bucketSize, unsigned bucketSizeFocus, float theta,
                                   const double* lim, const int* bnd);
cs_domain_t* cs_domain_create_u64f(int rank, int numRanks, unsigned bucketSize, unsigned bucketSizeFocus, float theta,
                                   const double* lim, const int* bnd);
cs_domain_t* cs_domain_create_u64d(int rank, int numRanks, unsigned bucketSize, unsigned bucketSizeFocus, float theta,
                                   const double* lim, const int* bnd);
void cs_domain_destroy(cs_domain_t* d);
int cs_domain_sync(cs_domain_t* d, const void* x, const void* y, const void* z, const void* h, const void* keys,
                   size_t n, int hostInput, void* stream);
/* out8 = {startIndex, endIndex, nParticlesWithHalos, numFocusLeaves, numFocusNodes, numGlobalLeaves,
 *         numGlobalNodes, maxTreeLevel}; box6 = current global box limits */
int cs_domain_info(const cs_domain_t* d, uint64_t* out8, double* box6);
void* cs_domain_ptr(cs_domain_t* d, int field);
int cs_domain_find_neighbors(cs_domain_t* d, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                             void* stream);
/* multi-rank domains (numRanks > 1): attach the communicator (same rank/size) before the first cs_domain_sync; the
 * communicator must outlive the domain */
int cs_domain_attach_comm(cs_domain_t* d, cs_comm_t* comm);
/* Domain::exchangeHalos(std::tie(fields...), sendBuf, recvBuf) (domain/domain.hpp:332-337): numArrays device arrays of
 * nParticlesWithHalos elements, elemBytes[k] bytes per element (multiples of 4); the halo elements of every array are
 * replaced by the owning ranks' values with the pattern recorded by the last cs_domain_sync.  No-op on one rank.
 * Synchronises the stream. */
int cs_domain_exchange_halos(cs_domain_t* d, void* const* arrays, const int* elemBytes, int numArrays, void* stream);
/* Domain::reapplySync(std::tie(fields...), sendBuf, recvBuf, ordering) (domain/domain.hpp:297-329) and the ExchangeLog
 * it replays (domain/index_ranges.hpp:186-210, GlobalAssignment::redoExchange, domain/assignment.hpp:206-215): move
 * further per-particle fields through the particle exchange and reordering of the LAST cs_domain_sync.  before[k] is
 * the field as it was when that sync was called (info4[0] elements, same order and layout as the coordinates the sync
 * consumed); after[k] (info4[1] = nParticlesWithHalos elements) receives the values of the assigned particles at
 * [startIndex, endIndex), halo elements are not written.  Collective over the ranks of the domain; synchronises the
 * stream.  cs_domain_replay_info: info4 = {elements before, elements after, startIndex, endIndex}. */
int cs_domain_reapply_sync(cs_domain_t* d, const void* const* before, void* const* after, const int* elemBytes,
                           int numArrays, void* stream);
int cs_domain_replay_info(const cs_domain_t* d, uint64_t* info4);
/* Domain::setHaloFactor (domain/domain.hpp:365): the halo search boxes of the following syncs use factor * 2h instead of
 * 2h (a deeper ghost layer for trees that are reused over several steps); multi-rank domains only, default 1 */
int cs_domain_set_halo_factor(cs_domain_t* d, float factor);
/* forget all tree state so that the next cs_domain_sync behaves like the first call on a new Domain (device buffers
 * are kept; used by bench.py to time cold syncs without re-allocating) */
int cs_domain_reset(cs_domain_t* d, void* stream);
/* device -> host copy of the synchronised arrays (NULL pointers are skipped); asynchronous on the stream */
int cs_domain_download(cs_domain_t* d, void* x, void* y, void* z, void* h, void* keys, void* stream);

/* markMacsGpu (traversal/collisions_gpu.h:62-71, macs.hpp:185-229): markings[i] = 1 for every node of the linked tree that
 * fails the MAC against one of the numFocusNodes leaves focusNodes[0..numFocusNodes] and is not contained in their key
 * range; centers4 = (x, y, z, mac^2) per node; limitSource != 0: nodes deeper than one level above the target leaf are
 * neither marked nor entered.  Marks are only set, never cleared. */
int cs_mark_macs_u32f(const uint32_t* prefixes, const int* childOffsets, const int* parents, const float* centers4,
                      const double* lim, const int* bnd, const uint32_t* focusNodes, int numFocusNodes, int limitSource,
                      uint8_t* markings, void* stream);
int cs_mark_macs_u64f(const uint64_t* prefixes, const int* childOffsets, const int* parents, const float* centers4,
                      const double* lim, const int* bnd, const uint64_t* focusNodes, int numFocusNodes, int limitSource,
                      uint8_t* markings, void* stream);
int cs_mark_macs_u64d(const uint64_t* prefixes, const int* childOffsets, const int* parents, const double* centers4,
                      const double* lim, const int* bnd, const uint64_t* focusNodes, int numFocusNodes, int limitSource,
                      uint8_t* markings, void* stream);
/* gatherRanges (halos/gather_halos_gpu.h:23-30): buffer[rangeScan[r] + k] = src[rangeOffsets[r] + k]; elements of
 * elemBytes bytes (a multiple of 4: int, util::array<float, 1..4>) */
int cs_gather_ranges(const uint32_t* rangeScan, const uint32_t* rangeOffsets, int numRanges, const void* src,
                     void* buffer, size_t bufferSize, int elemBytes, void* stream);
/* minMax (primitives/primitives_gpu.h:72-73): smallest and largest element of a device array, returned on the host
 * (synchronises the stream like the reference); n > 0 */
int cs_min_max_f(const float* first, size_t n, float* minOut, float* maxOut, void* stream);
int cs_min_max_d(const double* first, size_t n, double* minOut, double* maxOut, void* stream);
int cs_min_max_u32(const uint32_t* first, size_t n, uint32_t* minOut, uint32_t* maxOut, void* stream);

/* ---- target particle groups (traversal/groups_gpu.h:33-78) ----
 * computeFixedGroups(exec, first, last, groupSize, GroupData&): groups[g] = first + g * groupSize for the
 * ceil((last-first)/groupSize) groups, groups[numGroups] = last. */
int cs_compute_fixed_groups(uint32_t first, uint32_t last, uint32_t groupSize, uint32_t* groups, void* stream);
/* computeGroupSplits<Tc, T, KeyType>(exec, first, last, x, y, z, h, leaves, numLeaves, layout, box, groupSize, tolFactor,
 * numSplitsPerGroup, groups) for (double,double), (double,float), (float,float) and 64-bit keys: fixed groups of 32 or
 * 64 particles are cut wherever consecutive particles are further apart than min(tolFactor * cbrt(smallest leaf volume
 * of the group), 2h / minExtent) in unit-box coordinates.  In two calls, because the number of groups is only known on
 * the device: _begin returns it in *numGroupsOut (host memory; synchronises the stream, like the reference's read-back),
 * _finish then writes the numGroups + 1 ascending boundaries (groups[0] = first, groups[numGroups] = last).  Both calls
 * must come from the same host thread on the same stream, with nothing of this library in between on that stream. */
int cs_group_splits_begin_dd(uint32_t first, uint32_t last, const double* x, const double* y, const double* z,
                             const double* h, const uint64_t* leaves, int numLeaves, const uint32_t* layout,
                             const double* lim, const int* bnd, uint32_t groupSize, float tolFactor,
                             uint32_t* numGroupsOut, void* stream);
int cs_group_splits_begin_df(uint32_t first, uint32_t last, const double* x, const double* y, const double* z,
                             const float* h, const uint64_t* leaves, int numLeaves, const uint32_t* layout,
                             const double* lim, const int* bnd, uint32_t groupSize, float tolFactor,
                             uint32_t* numGroupsOut, void* stream);
int cs_group_splits_begin_ff(uint32_t first, uint32_t last, const float* x, const float* y, const float* z,
                             const float* h, const uint64_t* leaves, int numLeaves, const uint32_t* layout,
                             const double* lim, const int* bnd, uint32_t groupSize, float tolFactor,
                             uint32_t* numGroupsOut, void* stream);
int cs_group_splits_finish(uint32_t first, uint32_t last, uint32_t groupSize, uint32_t* groups, void* stream);

/* tuning hook: L2 fetch granularity hint in bytes (32, 64, 128) for the current device */
int cs_set_l2_fetch_granularity(int bytes);

/* experiment hook (not part of the drop-in surface): select a kernel variant, see csb::TuningKnob in csrc/common.cuh */
int cs_tuning_set(int knob, int value);

#ifdef __cplusplus
}
#endif

#endif /* CSTONE_B200_H */
