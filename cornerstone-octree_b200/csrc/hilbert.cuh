/* The reference's 3D Hilbert curve (sfc/hilbert.hpp:43-94,130-173) as a finite state machine.
 *
 * The reference walks the curve level by level: it maps the octant bits of the (already transformed) coordinates to a
 * key digit and then reflects / rotates the remaining low bits.  Composing those reflections and rotations gives a
 * transformation state (an axis permutation plus three flips) that depends only on the digits seen so far; the states
 * reachable from the identity are enumerated at compile time and the curve becomes two small tables:
 *     enc[state][raw octant bits]  -> (next state, key digit)      encode, most significant level first
 *     dec[state][key digit]        -> (next state, raw octant bits) decode, most significant level first
 * The bulk key kernel (sfc.cu) advances the same machine three levels per lookup; the kernels that work on node boxes
 * (centres, halo search boxes, MAC targets) use the one-level tables from shared memory.
 */
#pragma once

#include "common.cuh"

namespace csb
{

struct HilbertState
{
    int perm[3]; // transformed axis a reads raw axis perm[a] ...
    int flip[3]; // ... xor flip[a]
};

constexpr int hilbertStateCode(const HilbertState& s)
{
    return ((s.perm[0] * 3 + s.perm[1]) * 3 + s.perm[2]) * 8 + (s.flip[0] << 2 | s.flip[1] << 1 | s.flip[2]);
}

//! one level of the curve applied to the raw coordinate bits (rx,ry,rz) under transformation state s: the key digit of
//! this level and the state for the levels below (the octant-to-digit map and the reflect / rotate rules of
//! sfc/hilbert.hpp:59-91, composed onto s)
constexpr HilbertState hilbertStep(const HilbertState& s, unsigned rx, unsigned ry, unsigned rz, unsigned& digit)
{
    constexpr unsigned mortonToHilbert[8] = {0, 1, 3, 2, 7, 6, 4, 5};
    const unsigned raw[3]                 = {rx, ry, rz};
    unsigned t[3]                         = {0, 0, 0};
    for (int a = 0; a < 3; ++a)
        t[a] = raw[s.perm[a]] ^ unsigned(s.flip[a]);
    const unsigned xi = t[0], yi = t[1], zi = t[2];
    digit             = mortonToHilbert[(xi << 2) | (yi << 1) | zi];

    // which transformed axes get reflected below this level ...
    const unsigned F[3] = {xi & ((!yi) | zi), (xi & (yi | zi)) | (yi & (!zi)), (xi & (!yi) & (!zi)) | (yi & (!zi))};
    // ... and how they are rotated afterwards
    int q[3] = {0, 1, 2};
    if (zi) { q[0] = 1, q[1] = 2, q[2] = 0; }
    else if (!yi) { q[0] = 2, q[1] = 1, q[2] = 0; }

    HilbertState n{{0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < 3; ++a)
    {
        n.perm[a] = s.perm[q[a]];
        n.flip[a] = s.flip[q[a]] ^ int(F[q[a]] & 1u);
    }
    return n;
}

constexpr int hilbertMaxStates = 32;

struct HilbertFsm
{
    int numStates;
    HilbertState states[hilbertMaxStates];
    int idOf[6 * 27 * 8];
    unsigned char enc[hilbertMaxStates * 8]; // [state * 8 + raw octant] = next state << 3 | digit
    unsigned char dec[hilbertMaxStates * 8]; // [state * 8 + digit]      = next state << 3 | raw octant
};

constexpr HilbertFsm makeHilbertFsm()
{
    HilbertFsm f{};
    for (int& v : f.idOf)
        v = -1;
    f.states[0]                       = HilbertState{{0, 1, 2}, {0, 0, 0}};
    f.idOf[hilbertStateCode(f.states[0])] = 0;
    f.numStates                       = 1;
    for (int i = 0; i < f.numStates; ++i)
    {
        for (unsigned o = 0; o < 8; ++o)
        {
            unsigned d     = 0;
            HilbertState n = hilbertStep(f.states[i], (o >> 2) & 1, (o >> 1) & 1, o & 1, d);
            int code       = hilbertStateCode(n);
            if (f.idOf[code] < 0)
            {
                f.idOf[code]            = f.numStates;
                f.states[f.numStates++] = n;
            }
            f.enc[i * 8 + o] = static_cast<unsigned char>(f.idOf[code] << 3 | int(d));
            f.dec[i * 8 + d] = static_cast<unsigned char>(f.idOf[code] << 3 | int(o));
        }
    }
    return f;
}

inline constexpr HilbertFsm hilbertFsm = makeHilbertFsm();
static_assert(hilbertFsm.numStates <= hilbertMaxStates);

constexpr int hilbertTableBytes = 2 * hilbertMaxStates * 8;

#ifdef __CUDACC__
namespace detail
{
__device__ constexpr HilbertFsm hilbertFsmDevice = makeHilbertFsm();
}

//! copy the encode (first half) and decode (second half) tables into shared memory; all threads of the block call this,
//! followed by __syncthreads()
__device__ inline void stageHilbertTables(unsigned char* tables)
{
    for (int k = threadIdx.x; k < hilbertMaxStates * 8; k += blockDim.x)
    {
        tables[k]                        = detail::hilbertFsmDevice.enc[k];
        tables[hilbertMaxStates * 8 + k] = detail::hilbertFsmDevice.dec[k];
    }
}

//! Hilbert key of integer coordinates (iHilbert, sfc/hilbert.hpp:43-94)
template<class K>
__device__ inline K hilbertEncode(unsigned px, unsigned py, unsigned pz, const unsigned char* tables)
{
    K key          = 0;
    unsigned state = 0;
    for (int level = KeyTraits<K>::maxLevel - 1; level >= 0; --level)
    {
        unsigned octant = (((px >> level) & 1u) << 2) | (((py >> level) & 1u) << 1) | ((pz >> level) & 1u);
        unsigned e      = tables[state * 8 + octant];
        key             = (key << 3) | K(e & 7u);
        state           = e >> 3;
    }
    return key;
}

//! integer coordinates of a Hilbert key (decodeHilbert, sfc/hilbert.hpp:130-173)
template<class K>
__device__ inline void hilbertDecode(K key, unsigned& ox, unsigned& oy, unsigned& oz, const unsigned char* tables)
{
    const unsigned char* dec = tables + hilbertMaxStates * 8;
    unsigned x = 0, y = 0, z = 0, state = 0;
    for (int level = KeyTraits<K>::maxLevel - 1; level >= 0; --level)
    {
        unsigned digit = unsigned((key >> (3 * level)) & 7u);
        unsigned e     = dec[state * 8 + digit];
        x              = (x << 1) | ((e >> 2) & 1u);
        y              = (y << 1) | ((e >> 1) & 1u);
        z              = (z << 1) | (e & 1u);
        state          = e >> 3;
    }
    ox = x, oy = y, oz = z;
}

//! integer coordinates of a Morton key (sfc/morton.hpp:80-108): every third bit
template<class K>
__device__ inline void decodeMorton(K key, unsigned& ox, unsigned& oy, unsigned& oz)
{
    unsigned x = 0, y = 0, z = 0;
    for (unsigned b = 0; b < unsigned(KeyTraits<K>::maxLevel); ++b)
    {
        unsigned d = unsigned((key >> (3 * b)) & 7u);
        x |= ((d >> 2) & 1u) << b;
        y |= ((d >> 1) & 1u) << b;
        z |= (d & 1u) << b;
    }
    ox = x, oy = y, oz = z;
}
#endif

} // namespace csb
