/* per-level Hilbert / Morton integer encode and decode (sfc/hilbert.hpp:43-94,259-275, sfc/morton.hpp:80-108) for the
 * kernels that work on node boxes (centres, halo search boxes, MAC targets); the bulk key kernel in sfc.cu uses the
 * 3-levels-per-lookup table instead. */
#pragma once

#include "common.cuh"

namespace csb
{

//! per-level Hilbert encode (sfc/hilbert.hpp:43-94); only used for the two corner keys of a search box
template<class K>
__device__ inline K iHilbertLoop(unsigned px, unsigned py, unsigned pz)
{
    K key = 0;
    for (int level = KeyTraits<K>::maxLevel - 1; level >= 0; --level)
    {
        unsigned xi     = (px >> level) & 1u;
        unsigned yi     = (py >> level) & 1u;
        unsigned zi     = (pz >> level) & 1u;
        unsigned octant = (xi << 2) | (yi << 1) | zi;
        // mortonToHilbert = {0, 1, 3, 2, 7, 6, 4, 5} packed into one word, 3 bits per entry
        key = (key << 3) + K((0b101100110111010011001000u >> (3 * octant)) & 7u);

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

template<class K>
__device__ inline void decodeHilbert(K key, unsigned& ox, unsigned& oy, unsigned& oz)
{
    unsigned px = 0, py = 0, pz = 0;
    for (unsigned level = 0; level < unsigned(KeyTraits<K>::maxLevel); ++level)
    {
        unsigned octant = unsigned((key >> (3 * level)) & 7u);
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
    ox = px, oy = py, oz = pz;
}

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

} // namespace csb
