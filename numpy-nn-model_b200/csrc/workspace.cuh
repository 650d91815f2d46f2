// Bump allocator over the caller-provided workspace (the caller owns all scratch memory, like the
// reference's wrappers that cp.empty() every buffer; experimental/linear_swish/linear_swish_cutlass.py:157-171).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>

namespace nnb {

struct Bump {
    uint8_t* base;
    size_t cap;
    size_t off = 0;  // keeps growing past cap so callers can report the size they needed
    bool overflow = false;
    Bump(void* p, size_t bytes) : base(static_cast<uint8_t*>(p)), cap(p ? bytes : 0) {
        // align the base up to 256 bytes
        const uintptr_t a = reinterpret_cast<uintptr_t>(base);
        const uintptr_t al = (a + 255) & ~static_cast<uintptr_t>(255);
        const size_t skip = static_cast<size_t>(al - a);
        if (skip > cap) { cap = 0; } else { base += skip; cap -= skip; }
    }
    void* take(size_t bytes) {
        bytes = (bytes + 255) & ~static_cast<size_t>(255);
        void* r = (off + bytes <= cap) ? base + off : nullptr;
        if (!r && bytes > 0) overflow = true;
        off += bytes;
        return r;
    }
    size_t remaining() const { return off < cap ? ((cap - off) & ~static_cast<size_t>(255)) : 0; }
    bool ok() const { return !overflow; }
};

}  // namespace nnb
