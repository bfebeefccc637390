// Symmetric device heap: every device buffer of a handle is carved out of a few large cudaMalloc chunks by a
// deterministic first-fit allocator.  Ranks of a multi-GPU job create their handles with the same sizes and make the
// same allocation calls in the same order, so a buffer sits at the SAME (chunk, offset) on every rank; after the chunks
// have been exchanged as CUDA IPC handles (flip_dist_p2p_export / _import) the address of any local buffer on any peer
// is one addition away, and kernels can store straight into the other GPUs' copies over NVLink (xch.h).  This is the
// B200-shaped replacement for per-buffer halo staging: 180 GB per GPU holds the whole replicated state, NVSwitch makes
// every peer equally close, and nothing is ever packed or unpacked.
//
// The reference has no device memory at all; nothing here replaces reference code.
#pragma once
#include "rt.h"
#include <map>
#include <vector>

#define FLIP_MAX_RANKS 16
#define FLIP_HEAP_MAX_CHUNKS 8
#define FLIP_HEAP_ALIGN 512

struct SymHeap {
    struct Chunk {
        char *base = nullptr;
        size_t size = 0;
        std::map<size_t, size_t> free_;   // offset -> length of every free range (coalesced)
    };
    std::vector<Chunk> chunks;
    std::map<char *, size_t> live;        // allocation -> bytes
    size_t first_chunk = (size_t)64 << 20, grow_chunk = (size_t)64 << 20;
    bool frozen = false;                  // peers hold mappings of the current chunk list: it must not grow
    char *peer_base[FLIP_MAX_RANKS][FLIP_HEAP_MAX_CHUNKS] = {{nullptr}};
    int nranks = 1, rank = 0;

    static size_t round_up(size_t n) { return (n + FLIP_HEAP_ALIGN - 1) / FLIP_HEAP_ALIGN * FLIP_HEAP_ALIGN; }

    void *alloc(size_t bytes) {
        bytes = round_up(bytes ? bytes : 1);
        for (int pass = 0; pass < 2; pass++) {
            for (size_t c = 0; c < chunks.size(); c++) {
                Chunk &ch = chunks[c];
                for (auto it = ch.free_.begin(); it != ch.free_.end(); ++it) {
                    if (it->second < bytes) continue;
                    size_t off = it->first, len = it->second;
                    ch.free_.erase(it);
                    if (len > bytes) ch.free_[off + bytes] = len - bytes;
                    char *p = ch.base + off;
                    live[p] = bytes;
                    return p;
                }
            }
            if (pass == 1) break;
            if (frozen)
                throw FlipError("symmetric heap exhausted after the peer mappings were exchanged: load the scene (particles, "
                                "first solve) before flip_dist_p2p_export, or export/import again after the heap has grown");
            if (chunks.size() >= FLIP_HEAP_MAX_CHUNKS) throw FlipError("symmetric heap: too many chunks");
            size_t want = chunks.empty() ? first_chunk : grow_chunk;
            if (want < bytes) want = bytes;
            want = round_up(want);
            Chunk ch;
            cudaError_t e = cudaMalloc((void **)&ch.base, want);
            if (e != cudaSuccess) throw FlipError(std::string("symmetric heap: cudaMalloc of ") + std::to_string(want >> 20) + " MB failed: " + cudaGetErrorString(e));
            ch.size = want;
            ch.free_[0] = want;
            chunks.push_back(ch);
        }
        throw FlipError("symmetric heap: allocation failed");
    }

    void release(void *ptr) {
        if (!ptr) return;
        auto it = live.find((char *)ptr);
        if (it == live.end()) return;
        size_t bytes = it->second;
        live.erase(it);
        for (Chunk &ch : chunks) {
            if ((char *)ptr < ch.base || (char *)ptr >= ch.base + ch.size) continue;
            size_t off = (size_t)((char *)ptr - ch.base);
            auto nx = ch.free_.lower_bound(off);
            if (nx != ch.free_.end() && off + bytes == nx->first) { bytes += nx->second; nx = ch.free_.erase(nx); }
            if (nx != ch.free_.begin()) {
                auto pv = std::prev(nx);
                if (pv->first + pv->second == off) { pv->second += bytes; return; }
            }
            ch.free_[off] = bytes;
            return;
        }
    }

    // address of `local` in rank r's heap (valid once the peer mappings exist; r == rank gives `local` back)
    template <class T>
    T *peer(int r, T *local) const {
        if (!local || r == rank) return local;
        for (size_t c = 0; c < chunks.size(); c++) {
            const Chunk &ch = chunks[c];
            if ((char *)local >= ch.base && (char *)local < ch.base + ch.size) {
                if (!peer_base[r][c]) throw FlipError("symmetric heap: peer chunk is not mapped");
                return (T *)(peer_base[r][c] + ((char *)local - ch.base));
            }
        }
        throw FlipError("symmetric heap: pointer is not a heap allocation");
    }

    void destroy() {
        for (Chunk &ch : chunks) if (ch.base) cudaFree(ch.base);
        chunks.clear();
        live.clear();
        frozen = false;
    }
};
