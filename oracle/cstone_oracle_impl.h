/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the cornerstone domain-sync algorithms.
 * Included several times by cstone_oracle.c with different (KEY, REAL) type bindings.
 *
 * Required macros:  KEY (uint32_t|uint64_t), KBITS (32|64), MAXLEVEL (10|21), UNUSED_BITS (2|1),
 *                   KS(name)  -> name##_u32 | name##_u64          (key-only functions, emitted when ORC_EMIT_KEY)
 *                   REAL (float|double), RFLOOR, RCEIL, RRINT, RFABS, KT(name) -> name##_u32f | _u64f | _u64d
 * Every function cites the reference file:line (relative to /root/reference/include/cstone) it restates.
 * Parity status: pinned against the reference's golden vectors (tests/test_oracle_golden.py) and against
 * oracle/_ref/libcstone_ref.so, i.e. the unmodified reference headers (tests/test_oracle_vs_ref.py).
 */

#ifdef ORC_EMIT_KEY

#define NODE_RANGE0 ((KEY)1 << (3 * MAXLEVEL))

/* primitives/clz.hpp: countLeadingZeros, returns bit width for 0 */
static inline int KS(orc_clz)(KEY v)
{
    if (v == 0) return KBITS;
#if KBITS == 32
    return __builtin_clz(v);
#else
    return __builtin_clzll(v);
#endif
}

/* sfc/common.hpp:67-74 nodeRange */
static inline KEY KS(orc_node_range)(unsigned level) { return (KEY)1 << (3u * (MAXLEVEL - level)); }

/* sfc/common.hpp:120-125 treeLevel */
static inline unsigned KS(orc_tree_level)(KEY range) { return (KS(orc_clz)(range - 1) - UNUSED_BITS) / 3; }

/* sfc/common.hpp:112-116 commonPrefix */
static inline int KS(orc_common_prefix)(KEY a, KEY b) { return KS(orc_clz)(a ^ b) - UNUSED_BITS; }

/* sfc/common.hpp:172-180 encodePlaceholderBit */
static inline KEY KS(orc_encode_placeholder)(KEY code, int prefixLength)
{
    int nShifts = 3 * MAXLEVEL - prefixLength;
    return ((KEY)1 << prefixLength) | (code >> nShifts);
}

/* sfc/common.hpp:191-195 decodePrefixLength */
static inline unsigned KS(orc_decode_prefix_length)(KEY code) { return KBITS - 1 - KS(orc_clz)(code); }

/* sfc/common.hpp:198-206 decodePlaceholderBit */
static inline KEY KS(orc_decode_placeholder)(KEY code)
{
    int prefixLength = KS(orc_decode_prefix_length)(code);
    KEY ret          = code ^ ((KEY)1 << prefixLength);
    return ret << (3 * MAXLEVEL - prefixLength);
}

/* sfc/common.hpp:286-290 octalDigit */
static inline unsigned KS(orc_octal_digit)(KEY code, unsigned position)
{
    return (unsigned)((code >> (3u * (MAXLEVEL - position))) & 7u);
}

/* sfc/common.hpp:327-331 digitWeight */
static inline int KS(orc_digit_weight)(int digit)
{
    int fourGeqMask = -(int)(digit >= 4);
    return ((7 - digit) & fourGeqMask) - (digit & ~fourGeqMask);
}

/* sfc/hilbert.hpp:43-94 iHilbert */
KEY KS(orc_ihilbert)(unsigned px, unsigned py, unsigned pz)
{
    static const unsigned mortonToHilbert[8] = {0, 1, 3, 2, 7, 6, 4, 5};
    KEY key                                  = 0;
    for (int level = MAXLEVEL - 1; level >= 0; --level)
    {
        unsigned xi     = (px >> level) & 1u;
        unsigned yi     = (py >> level) & 1u;
        unsigned zi     = (pz >> level) & 1u;
        unsigned octant = (xi << 2) | (yi << 1) | zi;
        key             = (key << 3) + mortonToHilbert[octant];

        px ^= -(xi & ((!yi) | zi));
        py ^= -((xi & (yi | zi)) | (yi & (!zi)));
        pz ^= -((xi & (!yi) & (!zi)) | (yi & (!zi)));

        if (zi)
        {
            unsigned pt = px;
            px          = py;
            py          = pz;
            pz          = pt;
        }
        else if (!yi)
        {
            unsigned pt = px;
            px          = pz;
            pz          = pt;
        }
    }
    return key;
}

/* sfc/hilbert.hpp:130-173 decodeHilbert */
void KS(orc_decode_hilbert)(KEY key, unsigned* ox, unsigned* oy, unsigned* oz)
{
    unsigned px = 0, py = 0, pz = 0;
    for (unsigned level = 0; level < MAXLEVEL; ++level)
    {
        unsigned octant = (unsigned)((key >> (3 * level)) & 7u);
        unsigned xi     = octant >> 2u;
        unsigned yi     = (octant >> 1u) & 1u;
        unsigned zi     = octant & 1u;

        if (yi ^ zi)
        {
            unsigned pt = px;
            px          = pz;
            pz          = py;
            py          = pt;
        }
        else if ((!xi & !yi & !zi) || (xi & yi & zi))
        {
            unsigned pt = px;
            px          = pz;
            pz          = pt;
        }

        unsigned mask = (1u << level) - 1;
        px ^= mask & (-(xi & (yi | zi)));
        py ^= mask & (-((xi & ((!yi) | (!zi))) | ((!xi) & yi & zi)));
        pz ^= mask & (-((xi & (!yi) & (!zi)) | (yi & zi)));

        px |= (xi << level);
        py |= ((xi ^ yi) << level);
        pz |= ((yi ^ zi) << level);
    }
    *ox = px, *oy = py, *oz = pz;
}

/* sfc/morton.hpp:34-60 expandBits / :94-108 iMorton (bit interleave restated as a loop) */
KEY KS(orc_imorton)(unsigned ix, unsigned iy, unsigned iz)
{
    KEY key = 0;
    for (int b = MAXLEVEL - 1; b >= 0; --b)
    {
        key = (key << 3) | (KEY)((((ix >> b) & 1u) << 2) | (((iy >> b) & 1u) << 1) | ((iz >> b) & 1u));
    }
    return key;
}

/* sfc/morton.hpp:123-150 decodeMorton */
void KS(orc_decode_morton)(KEY key, unsigned* ox, unsigned* oy, unsigned* oz)
{
    unsigned x = 0, y = 0, z = 0;
    for (unsigned b = 0; b < MAXLEVEL; ++b)
    {
        unsigned d = (unsigned)((key >> (3 * b)) & 7u);
        x |= ((d >> 2) & 1u) << b;
        y |= ((d >> 1) & 1u) << b;
        z |= (d & 1u) << b;
    }
    *ox = x, *oy = y, *oz = z;
}

/* primitives/gather.hpp:42-67 sort_by_key: std::stable_sort of (key,value) pairs by key; restated as a
 * bottom-up stable merge sort */
void KS(orc_sort_by_key)(KEY* keys, uint32_t* values, size_t n)
{
    if (n < 2) return;
    KEY* kb      = (KEY*)malloc(n * sizeof(KEY));
    uint32_t* vb = (uint32_t*)malloc(n * sizeof(uint32_t));
    KEY *ks = keys, *kd = kb;
    uint32_t *vs = values, *vd = vb;
    for (size_t width = 1; width < n; width *= 2)
    {
#pragma omp parallel for schedule(dynamic, 64)
        for (size_t lo = 0; lo < n; lo += 2 * width)
        {
            size_t mid = lo + width < n ? lo + width : n;
            size_t hi  = lo + 2 * width < n ? lo + 2 * width : n;
            size_t i = lo, j = mid, o = lo;
            while (i < mid && j < hi)
            {
                if (ks[j] < ks[i]) { kd[o] = ks[j], vd[o] = vs[j], ++j; }
                else { kd[o] = ks[i], vd[o] = vs[i], ++i; }
                ++o;
            }
            while (i < mid)
                kd[o] = ks[i], vd[o] = vs[i], ++i, ++o;
            while (j < hi)
                kd[o] = ks[j], vd[o] = vs[j], ++j, ++o;
        }
        KEY* tk = ks;
        ks = kd, kd = tk;
        uint32_t* tv = vs;
        vs = vd, vd = tv;
    }
    if (ks != keys)
    {
        memcpy(keys, ks, n * sizeof(KEY));
        memcpy(values, vs, n * sizeof(uint32_t));
    }
    free(kb);
    free(vb);
}

static size_t KS(orc_lower_bound)(const KEY* a, size_t n, KEY v)
{
    size_t lo = 0, hi = n;
    while (lo < hi)
    {
        size_t m = lo + (hi - lo) / 2;
        if (a[m] < v) lo = m + 1;
        else hi = m;
    }
    return lo;
}

static size_t KS(orc_upper_bound)(const KEY* a, size_t n, KEY v)
{
    size_t lo = 0, hi = n;
    while (lo < hi)
    {
        size_t m = lo + (hi - lo) / 2;
        if (!(v < a[m])) lo = m + 1;
        else hi = m;
    }
    return lo;
}

/* tree/csarray.hpp:181-235 computeNodeCounts (+ :68-79 calculateNodeCount). The guess-narrowed search variant
 * (:99-168) returns the same lower bounds, so only the plain search is restated. */
void KS(orc_compute_node_counts)(
    const KEY* tree, unsigned* counts, int nNodes, const KEY* keys, size_t nKeys, unsigned maxCount)
{
    int firstNode = nNodes, lastNode = nNodes;
    if (nKeys)
    {
        firstNode = (int)KS(orc_upper_bound)(tree, nNodes, keys[0]) - 1;
        lastNode  = (int)KS(orc_upper_bound)(tree, nNodes, keys[nKeys - 1]);
    }
    for (int i = 0; i < firstNode; ++i)
        counts[i] = 0;
    for (int i = lastNode; i < nNodes; ++i)
        counts[i] = 0;
#pragma omp parallel for schedule(static)
    for (int i = firstNode; i < lastNode; ++i)
    {
        size_t a  = KS(orc_lower_bound)(keys, nKeys, tree[i]);
        size_t b  = KS(orc_lower_bound)(keys, nKeys, tree[i + 1]);
        size_t c  = b - a;
        counts[i] = (unsigned)(c < maxCount ? c : maxCount);
    }
}

/* tree/csarray.hpp:237-253 siblingAndLevel + :267-293 calculateNodeOp */
static int KS(orc_node_op)(const KEY* tree, int nodeIdx, const unsigned* counts, unsigned bucketSize)
{
    KEY thisNode   = tree[nodeIdx];
    KEY range      = tree[nodeIdx + 1] - thisNode;
    unsigned level = KS(orc_tree_level)(range);
    int siblingIdx = -1;
    if (level > 0)
    {
        siblingIdx = (int)KS(orc_octal_digit)(thisNode, level);
        int sib    = tree[nodeIdx - siblingIdx + 8] == tree[nodeIdx - siblingIdx] + KS(orc_node_range)(level - 1);
        if (!sib) siblingIdx = -1;
    }
    if (siblingIdx > 0)
    {
        const unsigned* g = counts + nodeIdx - siblingIdx;
        size_t parentCount =
            (size_t)g[0] + g[1] + (size_t)g[2] + g[3] + (size_t)g[4] + g[5] + (size_t)g[6] + (size_t)g[7];
        if (parentCount <= (size_t)bucketSize) return 0;
    }
    if (counts[nodeIdx] > bucketSize * 512 && level + 3 < MAXLEVEL) return 4096;
    if (counts[nodeIdx] > bucketSize * 64 && level + 2 < MAXLEVEL) return 512;
    if (counts[nodeIdx] > bucketSize * 8 && level + 1 < MAXLEVEL) return 64;
    if (counts[nodeIdx] > bucketSize && level < MAXLEVEL) return 8;
    return 1;
}

/* tree/csarray.hpp:306-329 rebalanceDecision */
int KS(orc_rebalance_decision)(const KEY* tree, const unsigned* counts, int nNodes, unsigned bucketSize, int* nodeOps)
{
    int converged = 1;
    for (int i = 0; i < nNodes; ++i)
    {
        int d = KS(orc_node_op)(tree, i, counts, bucketSize);
        if (d != 1) converged = 0;
        nodeOps[i] = d;
    }
    return converged;
}

/* tree/csarray.hpp:339-389 processNode + rebalanceTree. nodeOps has nNodes+1 entries (decisions in, scan out);
 * returns the new number of leaves; newTree must hold that + 1. Pass newTree==NULL to only scan. */
int KS(orc_rebalance_tree)(const KEY* tree, int nNodes, int* nodeOps, KEY* newTree)
{
    int sum = 0;
    for (int i = 0; i < nNodes; ++i)
    {
        int v      = nodeOps[i];
        nodeOps[i] = sum;
        sum += v;
    }
    nodeOps[nNodes] = sum;
    if (!newTree) return sum;
    for (int i = 0; i < nNodes; ++i)
    {
        KEY thisNode   = tree[i];
        unsigned level = KS(orc_tree_level)(tree[i + 1] - thisNode);
        int opCode     = nodeOps[i + 1] - nodeOps[i];
        int at         = nodeOps[i];
        if (opCode == 1) newTree[at] = thisNode;
        else if (opCode >= 8)
        {
            unsigned levelDiff = opCode == 8 ? 1 : opCode == 64 ? 2 : opCode == 512 ? 3 : 4;
            for (int s = 0; s < opCode; ++s)
                newTree[at + s] = thisNode + (KEY)s * KS(orc_node_range)(level + levelDiff);
        }
    }
    newTree[sum] = tree[nNodes];
    return sum;
}

/* tree/csarray.hpp:408-426 updateOctree: leaves/counts are updated in place (capacity cap leaves);
 * returns new leaf count or -needed if cap is too small */
long KS(orc_update_octree)(
    const KEY* keys, size_t n, unsigned bucket, KEY* leaves, unsigned* counts, long nLeaves, long cap, int* converged)
{
    int* ops    = (int*)malloc((nLeaves + 1) * sizeof(int));
    *converged  = KS(orc_rebalance_decision)(leaves, counts, (int)nLeaves, bucket, ops);
    ops[nLeaves] = 0;
    int* scan   = (int*)malloc((nLeaves + 1) * sizeof(int));
    memcpy(scan, ops, (nLeaves + 1) * sizeof(int));
    long newN = KS(orc_rebalance_tree)(leaves, (int)nLeaves, scan, NULL);
    if (newN > cap)
    {
        free(ops), free(scan);
        return -newN;
    }
    KEY* nt = (KEY*)malloc((newN + 1) * sizeof(KEY));
    KS(orc_rebalance_tree)(leaves, (int)nLeaves, ops, nt);
    memcpy(leaves, nt, (newN + 1) * sizeof(KEY));
    KS(orc_compute_node_counts)(leaves, counts, (int)newN, keys, n, 0xFFFFFFFFu);
    free(ops), free(scan), free(nt);
    return newN;
}

/* tree/csarray.hpp:429-440 computeOctree */
long KS(orc_compute_octree)(const KEY* keys, size_t n, unsigned bucket, KEY* leaves, unsigned* counts, long cap)
{
    if (cap < 1) return -1;
    leaves[0] = 0, leaves[1] = NODE_RANGE0;
    counts[0]   = (unsigned)n;
    long nl     = 1;
    int conv    = 0;
    while (!conv)
    {
        nl = KS(orc_update_octree)(keys, n, bucket, leaves, counts, nl, cap, &conv);
        if (nl < 0) return nl;
    }
    return nl;
}

/* tree/octree.hpp:40-50 binaryKeyWeight */
static int KS(orc_binary_key_weight)(KEY key, unsigned level)
{
    int ret = 0;
    for (unsigned l = 1; l <= level + 1; ++l)
        ret += KS(orc_digit_weight)((int)KS(orc_octal_digit)(key, l));
    return ret;
}

/* tree/octree.hpp:52-196 buildOctreeCpu = createUnsortedLayoutCpu + sort_by_key + invert + getLevelRangeCpu +
 * linkTreeCpu.  childOffsets: numNodes entries, parents: (numNodes-1)/8, levelRange: MAXLEVEL+2 */
void KS(orc_build_octree)(const KEY* leaves,
                          int numLeaves,
                          KEY* prefixes,
                          int* childOffsets,
                          int* parents,
                          int* levelRange,
                          int* internalToLeaf,
                          int* leafToInternal)
{
    int numInternal = (numLeaves - 1) / 7;
    int numNodes    = numLeaves + numInternal;
    for (int tid = 0; tid < numLeaves; ++tid)
    {
        KEY key                        = leaves[tid];
        unsigned level                 = KS(orc_tree_level)(leaves[tid + 1] - key);
        prefixes[tid + numInternal]    = KS(orc_encode_placeholder)(key, 3 * level);
        internalToLeaf[tid + numInternal] = tid + numInternal;
        unsigned prefixLength          = KS(orc_common_prefix)(key, leaves[tid + 1]);
        if (prefixLength % 3 == 0 && tid < numLeaves - 1)
        {
            int octIndex             = (tid + KS(orc_binary_key_weight)(key, prefixLength / 3)) / 7;
            prefixes[octIndex]       = KS(orc_encode_placeholder)(key, prefixLength);
            internalToLeaf[octIndex] = octIndex;
        }
    }
    KS(orc_sort_by_key)(prefixes, (uint32_t*)internalToLeaf, numNodes);
    for (int i = 0; i < numNodes; ++i)
        leafToInternal[internalToLeaf[i]] = i;
    for (int i = 0; i < numNodes; ++i)
        internalToLeaf[i] -= numInternal;
    for (unsigned level = 0; level <= MAXLEVEL; ++level)
        levelRange[level] = (int)KS(orc_lower_bound)(prefixes, numNodes, KS(orc_encode_placeholder)(0, 3 * level));
    levelRange[MAXLEVEL + 1] = numNodes;

    for (int i = 0; i < numNodes; ++i)
        childOffsets[i] = 0;
    for (int i = 0; i < numInternal; ++i)
    {
        int idxA              = leafToInternal[i];
        KEY prefix            = prefixes[idxA];
        KEY nodeKey           = KS(orc_decode_placeholder)(prefix);
        unsigned prefixLength = KS(orc_decode_prefix_length)(prefix);
        unsigned level        = prefixLength / 3;
        KEY childPrefix       = KS(orc_encode_placeholder)(nodeKey, prefixLength + 3);
        int s0                = levelRange[level + 1];
        int s1                = levelRange[level + 2];
        int childIdx          = s0 + (int)KS(orc_lower_bound)(prefixes + s0, s1 - s0, childPrefix);
        if (childIdx != s1 && childPrefix == prefixes[childIdx])
        {
            childOffsets[idxA]          = childIdx;
            parents[(childIdx - 1) / 8] = idxA;
        }
    }
}

/* tree/octree.hpp:572-615 upsweep with NodeCount (sum of 8 children capped at 2^32-1) */
void KS(orc_upsweep_counts)(const int* levelRange, const int* childOffsets, unsigned* counts)
{
    for (int level = MAXLEVEL; level >= 0; --level)
    {
        for (int i = levelRange[level]; i < levelRange[level + 1]; ++i)
        {
            int c = childOffsets[i];
            if (c)
            {
                uint64_t sum = 0;
                for (int o = 0; o < 8; ++o)
                    sum += counts[c + o];
                counts[i] = (unsigned)(sum < 0xFFFFFFFFull ? sum : 0xFFFFFFFFull);
            }
        }
    }
}

/* traversal/traversal.hpp:26-69 singleTraversal, restated with callbacks */
typedef int (*KS(orc_continue_fn))(int node, void* ctx);
typedef void (*KS(orc_leaf_fn))(int node, void* ctx);
static void KS(orc_single_traversal)(
    const int* childOffsets, const int* parents, KS(orc_continue_fn) cont, KS(orc_leaf_fn) leafAction, void* ctx)
{
    if (!cont(0, ctx)) return;
    if (childOffsets[0] == 0)
    {
        if (leafAction) leafAction(0, ctx);
        return;
    }
    int node      = childOffsets[0];
    int backtrack = 0;
    while (node != 0)
    {
        int isLeaf  = childOffsets[node] == 0;
        int descend = !backtrack && cont(node, ctx);
        if (isLeaf && descend && leafAction) leafAction(node, ctx);
        int siblingIdx = (node - 1) % 8;
        if (!isLeaf && descend)
        {
            node      = childOffsets[node];
            backtrack = 0;
        }
        else if (siblingIdx < 7)
        {
            node++;
            backtrack = 0;
        }
        else
        {
            node      = parents[(node - 1) / 8];
            backtrack = 1;
        }
    }
}


/* ---------------------------------------------------------------- focus-tree (LET) rebalance decisions */

/* sfc/common.hpp:369-386 lastNzPlace, makePrefix */
static inline int KS(orc_last_nz_place)(KEY x)
{
    if (!x) return MAXLEVEL;
    int ctz = 0;
    while (((x >> ctz) & 1u) == 0)
        ++ctz;
    return MAXLEVEL - ctz / 3;
}

static inline KEY KS(orc_make_prefix)(KEY a)
{
    if (a == 0) return 1;
    return KS(orc_encode_placeholder)(a, 3 * KS(orc_last_nz_place)(a));
}

/* tree/octree.hpp:198-216 containingNode */
static int KS(orc_containing_node)(KEY nodeKey, const KEY* prefixes, const int* childOffsets)
{
    int nodeLevel = (int)(KS(orc_decode_prefix_length)(nodeKey) / 3);
    KEY key       = KS(orc_decode_placeholder)(nodeKey);
    int ret       = 0;
    for (int i = 1; i <= nodeLevel; ++i)
    {
        if (childOffsets[ret] == 0 || nodeKey == prefixes[ret]) break;
        ret = childOffsets[ret] + (int)KS(orc_octal_digit)(key, (unsigned)i);
    }
    return ret;
}

/* focus/rebalance.hpp:31-74 mergeCountAndMacOp (overlapTwoRanges: traversal/boxoverlap.hpp:26-30) */
static int KS(orc_merge_count_and_mac_op)(int nodeIdx, const KEY* nodeKeys, const int* childOffsets, const int* parents,
                                          const unsigned* counts, const uint8_t* macs, KEY focusStart, KEY focusEnd,
                                          unsigned bucketSize)
{
    int siblingGroup = (nodeIdx - 1) / 8;
    int parent       = nodeIdx ? parents[siblingGroup] : 0;
    KEY nodeKey      = nodeKeys[nodeIdx];
    unsigned level   = KS(orc_decode_prefix_length)(nodeKey) / 3;
    if (nodeIdx)
    {
        int countMerge    = counts[parent] <= bucketSize;
        int macMerge      = macs[parent] == 0;
        KEY firstGroupKey = KS(orc_decode_placeholder)(nodeKeys[parent]);
        KEY lastGroupKey  = firstGroupKey + 8 * KS(orc_node_range)(level);
        int inFringe      = lastGroupKey > focusStart && focusEnd > firstGroupKey;
        if (countMerge || (macMerge && !inFringe)) return 0;
    }
    KEY nodeStart = KS(orc_decode_placeholder)(nodeKey);
    int isLeaf    = childOffsets[nodeIdx] == 0;
    int inFocus   = nodeStart >= focusStart && nodeStart < focusEnd;
    if (isLeaf && (macs[nodeIdx] || inFocus))
    {
        if (level + 3 < MAXLEVEL && counts[nodeIdx] > 4096 * bucketSize) return 4096;
        if (level + 2 < MAXLEVEL && counts[nodeIdx] > 512 * bucketSize) return 512;
        if (level + 1 < MAXLEVEL && counts[nodeIdx] > 64 * bucketSize) return 64;
        if (level < MAXLEVEL && counts[nodeIdx] > bucketSize) return 8;
    }
    return 1;
}

/* focus/rebalance.hpp:138-156 rebalanceDecisionEssential */
void KS(orc_rebalance_decision_essential)(const KEY* nodeKeys, int numNodes, const int* childOffsets, const int* parents,
                                          const unsigned* counts, const uint8_t* macs, KEY focusStart, KEY focusEnd,
                                          unsigned bucketSize, int* nodeOps)
{
    for (int i = 0; i < numNodes; ++i)
        nodeOps[i] = KS(orc_merge_count_and_mac_op)(i, nodeKeys, childOffsets, parents, counts, macs, focusStart,
                                                    focusEnd, bucketSize);
}

/* focus/rebalance.hpp:91-116 nzAncestorOp, :158-172 protectAncestors (in place, ascending node index - parents have
 * smaller indices than their children, so the serial order is one of the orders the reference's parallel loop allows) */
int KS(orc_protect_ancestors)(const KEY* nodeKeys, int numNodes, const int* parents, int* nodeOps)
{
    int numChanges = 0;
    for (int i = 0; i < numNodes; ++i)
    {
        int decision;
        if (i == 0) decision = nodeOps[0];
        else
        {
            int a = i;
            while (nodeOps[a] == 0)
                a = parents[(a - 1) / 8];
            decision = KS(orc_decode_placeholder)(nodeKeys[i]) == KS(orc_decode_placeholder)(nodeKeys[a]) ? nodeOps[a] : 0;
        }
        if (decision != 1) numChanges++;
        nodeOps[i] = decision;
    }
    return numChanges == 0;
}

/* focus/rebalance.hpp:174-252 enforceKeySingle + enforceKeys; returns the ResolutionStatus
 * (0 converged, 1 cancelMerge, 2 rebalance, 3 failed) */
int KS(orc_enforce_keys)(const KEY* mandatoryKeys, int numKeys, const KEY* nodeKeys, const int* childOffsets,
                         const int* parents, int* nodeOps)
{
    int status = 0;
    for (int k = 0; k < numKeys; ++k)
    {
        KEY key = mandatoryKeys[k];
        if (key == 0 || key == NODE_RANGE0) continue;
        int st            = 0;
        KEY nodeKeyWant   = KS(orc_make_prefix)(key);
        int nodeIdx       = KS(orc_containing_node)(nodeKeyWant, nodeKeys, childOffsets);
        KEY nodeKeyHave   = nodeKeys[nodeIdx];
        int nodeLevelHave = (int)(KS(orc_decode_prefix_length)(nodeKeyHave) / 3);
        int trySplit      = nodeKeyHave != nodeKeyWant && nodeLevelHave < MAXLEVEL;
        int undoMerges    = nodeOps[nodeIdx] == 0 || trySplit;
        if (undoMerges && nodeIdx > 0)
        {
            st         = 1;
            int parent = nodeIdx;
            do
            {
                parent           = parents[(parent - 1) / 8];
                int firstSibling = childOffsets[parent];
                for (int i = firstSibling; i < firstSibling + 8; ++i)
                    if (nodeOps[i] == 0) nodeOps[i] = 1;
            } while (parent != 0);
        }
        if (trySplit)
        {
            int levelDiff = KS(orc_last_nz_place)(key) - nodeLevelHave;
            st            = levelDiff > 1 ? 3 : 2;
            if (levelDiff > 1) levelDiff = 1;
            int op = 1 << (3 * levelDiff);
            if (op > nodeOps[nodeIdx]) nodeOps[nodeIdx] = op;
        }
        if (st > status) status = st;
    }
    return status;
}

#endif /* ORC_EMIT_KEY */

/* ====================================================================================================== */
#ifdef ORC_EMIT_KEY_REAL

/* sfc/box.hpp:100-122 Box: limits, lengths and inverse lengths are all stored in REAL */
typedef struct
{
    REAL lim[6];
    REAL len[3];
    REAL ilen[3];
    int bnd[3];
} KT(OrcBox);

static KT(OrcBox) KT(orc_make_box)(const double* lim, const int* bnd)
{
    KT(OrcBox) b;
    for (int i = 0; i < 6; ++i)
        b.lim[i] = (REAL)lim[i];
    for (int d = 0; d < 3; ++d)
    {
        b.len[d]  = b.lim[2 * d + 1] - b.lim[2 * d];
        b.ilen[d] = (REAL)1 / (b.lim[2 * d + 1] - b.lim[2 * d]);
        b.bnd[d]  = bnd[d];
    }
    return b;
}

/* sfc/sfc.hpp:142-179 sfc3D */
static KEY KT(orc_sfc3d)(int kind, REAL x, REAL y, REAL z, const KT(OrcBox) * box)
{
    const unsigned cubeLength = 1u << MAXLEVEL;
    const int mcoord          = (int)(cubeLength - 1);
    REAL mx = cubeLength * box->ilen[0], my = cubeLength * box->ilen[1], mz = cubeLength * box->ilen[2];
    int ix = (int)(RFLOOR(x * mx) - box->lim[0] * mx);
    int iy = (int)(RFLOOR(y * my) - box->lim[2] * my);
    int iz = (int)(RFLOOR(z * mz) - box->lim[4] * mz);
    ix     = ix < mcoord ? ix : mcoord;
    iy     = iy < mcoord ? iy : mcoord;
    iz     = iz < mcoord ? iz : mcoord;
    return kind == 0 ? KS(orc_ihilbert)(ix, iy, iz) : KS(orc_imorton)(ix, iy, iz);
}

/* sfc/sfc.hpp:268-283 computeSfcKeys (keys equal to removeKey = 2^(3 maxLevel) are left alone) */
void KT(orc_sfc_keys)(
    int kind, const REAL* x, const REAL* y, const REAL* z, KEY* keys, size_t n, const double* lim, const int* bnd)
{
    KT(OrcBox) box = KT(orc_make_box)(lim, bnd);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i)
    {
        if (keys[i] != NODE_RANGE0) keys[i] = KT(orc_sfc3d)(kind, x[i], y[i], z[i], &box);
    }
}

/* focus/source_center.hpp:146-158 nodeFpCenters + sfc/hilbert.hpp:259-275 hilbertIBox / morton.hpp:160-167 +
 * sfc/box.hpp:318-335 centerAndSize.  kind 0 = Hilbert (what the reference hard-codes through SfcKind). */
void KT(orc_node_fp_centers)(
    int kind, const KEY* prefixes, size_t n, REAL* centers, REAL* sizes, const double* lim, const int* bnd)
{
    KT(OrcBox) box = KT(orc_make_box)(lim, bnd);
    const int maxCoord = 1 << MAXLEVEL;
    const REAL uL      = (REAL)1 / maxCoord;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i)
    {
        KEY prefix          = prefixes[i];
        KEY startKey        = KS(orc_decode_placeholder)(prefix);
        unsigned level      = KS(orc_decode_prefix_length)(prefix) / 3;
        unsigned cubeLength = (unsigned)maxCoord >> level;
        unsigned mask       = ~(cubeLength - 1);
        unsigned ix, iy, iz;
        if (kind == 0) KS(orc_decode_hilbert)(startKey, &ix, &iy, &iz);
        else KS(orc_decode_morton)(startKey, &ix, &iy, &iz);
        ix &= mask, iy &= mask, iz &= mask;
        int imin[3] = {(int)ix, (int)iy, (int)iz};
        for (int d = 0; d < 3; ++d)
        {
            int imax           = imin[d] + (int)cubeLength;
            REAL halfUnit      = (REAL)0.5 * uL * box.len[d];
            centers[3 * i + d] = box.lim[2 * d] + (imax + imin[d]) * halfUnit;
            sizes[3 * i + d]   = (imax - imin[d]) * halfUnit;
        }
    }
}

/* focus/source_center.hpp:28-43 computeBoundingBox over leaves [firstLeaf,lastLeaf); searchCenters[leaf] holds the
 * init point (geometric centre) on entry, as in focus/octree_focus_mpi.hpp:539-557 */
void KT(orc_bounding_boxes)(const REAL* x,
                            const REAL* y,
                            const REAL* z,
                            const REAL* h,
                            const uint32_t* layout,
                            int firstLeaf,
                            int lastLeaf,
                            REAL scale,
                            REAL* sc,
                            REAL* ss)
{
#pragma omp parallel for schedule(static)
    for (int l = firstLeaf; l < lastLeaf; ++l)
    {
        REAL mn[3] = {sc[3 * l], sc[3 * l + 1], sc[3 * l + 2]};
        REAL mx[3] = {mn[0], mn[1], mn[2]};
        for (uint32_t i = layout[l]; i < layout[l + 1]; ++i)
        {
            REAL r    = h[i] * scale;
            REAL p[3] = {x[i], y[i], z[i]};
            for (int d = 0; d < 3; ++d)
            {
                REAL lo = p[d] - r, hi = p[d] + r;
                mn[d] = lo < mn[d] ? lo : mn[d];
                mx[d] = hi > mx[d] ? hi : mx[d];
            }
        }
        for (int d = 0; d < 3; ++d)
        {
            sc[3 * l + d] = (mx[d] + mn[d]) * (REAL)0.5;
            ss[3 * l + d] = (mx[d] - mn[d]) * (REAL)0.5;
        }
    }
}

/* sfc/box.hpp:176-190 applyPbc on one component */
static inline REAL KT(orc_pbc_fold)(REAL d, int dim, const KT(OrcBox) * box)
{
    int pbc = box->bnd[dim] == 1;
    return d - pbc * box->len[dim] * RRINT(d * box->ilen[dim]);
}

/* traversal/boxoverlap.hpp:280-291 overlap(aCenter,aSize,bCenter,bSize,box) */
static int KT(orc_overlap)(const REAL* ac, const REAL* as, const REAL* bc, const REAL* bs, const KT(OrcBox) * box)
{
    for (int d = 0; d < 3; ++d)
    {
        REAL dx = bc[d] - ac[d];
        dx      = RFABS(KT(orc_pbc_fold)(dx, d, box));
        dx -= as[d];
        dx -= bs[d];
        if (!(dx < (REAL)0)) return 0;
    }
    return 1;
}

/* traversal/boxoverlap.hpp:127-152 containedIn(codeStart, codeEnd, center, size, box).
 * Always Hilbert-encodes (SfcKind) as in the reference. */
static int KT(orc_contained_in)(KEY codeStart, KEY codeEnd, const REAL* c, const REAL* s, const KT(OrcBox) * box)
{
    REAL bmin[3], bmax[3];
    REAL dFromMin = 0, dFromMax = 0;
    for (int d = 0; d < 3; ++d)
    {
        bmin[d] = c[d] - s[d];
        bmax[d] = c[d] + s[d];
        REAL a  = bmin[d] - box->lim[2 * d];
        REAL b  = bmax[d] - box->lim[2 * d + 1];
        if (d == 0) dFromMin = a, dFromMax = b;
        else
        {
            dFromMin = a < dFromMin ? a : dFromMin;
            dFromMax = b > dFromMax ? b : dFromMax;
        }
    }
    if (dFromMin < (REAL)0 || dFromMax > (REAL)0) return codeStart == 0 && codeEnd == NODE_RANGE0;

    const int gridDim = 1 << MAXLEVEL;
    for (int d = 0; d < 3; ++d)
        bmax[d] += box->len[d] * ((REAL)1 / gridDim);

    KEY lowCode      = KT(orc_sfc3d)(0, bmin[0], bmin[1], bmin[2], box);
    KEY highCode     = KT(orc_sfc3d)(0, bmax[0], bmax[1], bmax[2], box);
    unsigned level   = KS(orc_common_prefix)(lowCode, highCode) / 3;
    KEY nodeStart    = lowCode & ~(KS(orc_node_range)(level) - 1);
    KEY nodeEnd      = nodeStart + KS(orc_node_range)(level);
    return nodeStart >= codeStart && nodeEnd <= codeEnd;
}

typedef struct
{
    const KEY* prefixes;
    const REAL* centers;
    const REAL* sizes;
    const REAL* tc;
    const REAL* ts;
    const KT(OrcBox) * box;
    KEY exStart, exEnd;
    uint8_t* flags;
} KT(OrcHaloCtx);

/* traversal/collisions.hpp:25-51 findCollisions continuation */
static int KT(orc_halo_continue)(int idx, void* vctx)
{
    KT(OrcHaloCtx)* c   = (KT(OrcHaloCtx)*)vctx;
    KEY prefix          = c->prefixes[idx];
    unsigned prefixLen  = KS(orc_decode_prefix_length)(prefix);
    KEY nk1             = KS(orc_decode_placeholder)(prefix);
    KEY nk2             = nk1 + ((KEY)1 << (3 * MAXLEVEL - prefixLen));
    int contained       = !(nk1 < c->exStart || nk2 > c->exEnd);
    int ov = !contained && KT(orc_overlap)(c->centers + 3 * idx, c->sizes + 3 * idx, c->tc, c->ts, c->box);
    if (ov) c->flags[idx] = 1;
    return ov;
}

/* traversal/collisions.hpp:53-94 findHalos */
void KT(orc_find_halos)(const KEY* prefixes,
                        const int* childOffsets,
                        const int* parents,
                        const REAL* centers,
                        const REAL* sizes,
                        const KEY* leaves,
                        const REAL* searchCenters,
                        const REAL* searchSizes,
                        const double* lim,
                        const int* bnd,
                        int firstNode,
                        int lastNode,
                        uint8_t* flags)
{
    KT(OrcBox) box = KT(orc_make_box)(lim, bnd);
    KEY lowestKey = leaves[firstNode], highestKey = leaves[lastNode];
    for (int l = firstNode; l < lastNode; ++l)
    {
        if (KT(orc_contained_in)(lowestKey, highestKey, searchCenters + 3 * l, searchSizes + 3 * l, &box)) continue;
        KT(OrcHaloCtx) ctx = {prefixes, centers,   sizes,      searchCenters + 3 * l, searchSizes + 3 * l,
                              &box,     lowestKey, highestKey, flags};
        KS(orc_single_traversal)(childOffsets, parents, KT(orc_halo_continue), NULL, &ctx);
    }
}

typedef struct
{
    uint32_t i;
    REAL p[3];
    REAL radiusSq, cellRadiusSq;
    int usePbc;
    const REAL *x, *y, *z;
    const REAL* centers;
    const REAL* sizes;
    const int* internalToLeaf;
    const uint32_t* layout;
    const KT(OrcBox) * box;
    unsigned ngmax, numNeighbors;
    uint32_t* neighbors;
} KT(OrcNbCtx);

/* findneighbors.hpp:108-112 overlaps/overlapsPbc with traversal/boxoverlap.hpp:229-250 minDistance */
static int KT(orc_nb_continue)(int idx, void* vctx)
{
    KT(OrcNbCtx)* c = (KT(OrcNbCtx)*)vctx;
    REAL sq[3];
    for (int d = 0; d < 3; ++d)
    {
        REAL dx;
        if (c->usePbc)
        {
            dx = c->centers[3 * idx + d] - c->p[d];
            dx = RFABS(KT(orc_pbc_fold)(dx, d, c->box));
            dx -= c->sizes[3 * idx + d];
        }
        else { dx = RFABS(c->centers[3 * idx + d] - c->p[d]) - c->sizes[3 * idx + d]; }
        dx += RFABS(dx);
        dx *= (REAL)0.5;
        sq[d] = dx * dx;
    }
    /* util/array.hpp:236-240,316-320 norm2 = dot(a,a) = ((a[Is]*a[Is]) + ...), a unary RIGHT fold: x*x + (y*y + z*z) */
    REAL n2 = sq[0] + (sq[1] + sq[2]);
    return n2 < c->cellRadiusSq;
}

/* findneighbors.hpp:114-148 searchBox/searchBoxPbc with :33-60 distanceSq */
static void KT(orc_nb_leaf)(int idx, void* vctx)
{
    KT(OrcNbCtx)* c = (KT(OrcNbCtx)*)vctx;
    int leafIdx    = c->internalToLeaf[idx];
    for (uint32_t j = c->layout[leafIdx]; j < c->layout[leafIdx + 1]; ++j)
    {
        if (j == c->i) continue;
        REAL dx = c->x[j] - c->p[0], dy = c->y[j] - c->p[1], dz = c->z[j] - c->p[2];
        if (c->usePbc)
        {
            dx = KT(orc_pbc_fold)(dx, 0, c->box);
            dy = KT(orc_pbc_fold)(dy, 1, c->box);
            dz = KT(orc_pbc_fold)(dz, 2, c->box);
        }
        REAL d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < c->radiusSq)
        {
            if (c->numNeighbors < c->ngmax) c->neighbors[c->numNeighbors] = j;
            c->numNeighbors++;
        }
    }
}

/* findneighbors.hpp:77-177 findNeighbors over [first,last): neighbors[(i-first)*ngmax + k], neighborsCount[i-first]
 * (count is not truncated, list is). h has the coordinate type here (Th == Tc). */
void KT(orc_find_neighbors)(const REAL* x,
                            const REAL* y,
                            const REAL* z,
                            const REAL* h,
                            uint32_t first,
                            uint32_t last,
                            const double* lim,
                            const int* bnd,
                            const int* childOffsets,
                            const int* parents,
                            const int* internalToLeaf,
                            const uint32_t* layout,
                            const REAL* centers,
                            const REAL* sizes,
                            unsigned ngmax,
                            uint32_t* neighbors,
                            unsigned* neighborsCount)
{
    KT(OrcBox) box = KT(orc_make_box)(lim, bnd);
    int anyPbc     = box.bnd[0] == 1 || box.bnd[1] == 1 || box.bnd[2] == 1;
#pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t k = 0; k < last - first; ++k)
    {
        uint32_t i = k + first;
        KT(OrcNbCtx) c;
        c.i = i;
        c.p[0] = x[i], c.p[1] = y[i], c.p[2] = z[i];
        REAL hi        = h[i];
        c.radiusSq     = (REAL)4.0 * hi * hi;
        c.cellRadiusSq = c.radiusSq * 1.0f * 1.0f; /* searchExtFactor = 1 */
        int inside     = 1;
        for (int d = 0; d < 3; ++d)
        {
            REAL s = (REAL)2 * hi;
            if (!(c.p[d] - s >= box.lim[2 * d] && c.p[d] + s <= box.lim[2 * d + 1])) inside = 0;
        }
        c.usePbc = anyPbc && !inside;
        c.x = x, c.y = y, c.z = z;
        c.centers = centers, c.sizes = sizes;
        c.internalToLeaf = internalToLeaf, c.layout = layout;
        c.box            = &box;
        c.ngmax = ngmax, c.numNeighbors = 0;
        c.neighbors = neighbors + (size_t)k * ngmax;
        KS(orc_single_traversal)(childOffsets, parents, KT(orc_nb_continue), KT(orc_nb_leaf), &c);
        neighborsCount[k] = c.numNeighbors;
    }
}

#endif /* ORC_EMIT_KEY_REAL */
