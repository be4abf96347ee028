// Shim: the subset of boost::dynamic_bitset<> used by the reference's AD-tree and prune code.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
namespace boost {
template <class Block = unsigned long, class Alloc = std::allocator<Block>>
class dynamic_bitset {
public:
    typedef std::size_t size_type;
    static const size_type npos = static_cast<size_type>(-1);
    dynamic_bitset() : nbits(0) {}
    explicit dynamic_bitset(size_type n, unsigned long value = 0) : w((n + 63) / 64, 0), nbits(n) {
        if (n && value) { w[0] = value; trim(); }
    }
    size_type size() const { return nbits; }
    size_type num_blocks() const { return w.size(); }
    void resize(size_type n, bool value = false) {
        size_type old = nbits;
        w.resize((n + 63) / 64, value ? ~0ULL : 0ULL);
        nbits = n;
        if (value) for (size_type i = old; i < n && i < ((old + 63) / 64) * 64; i++) set(i);
        trim();
    }
    void clear() { w.clear(); nbits = 0; }
    dynamic_bitset &set() { for (auto &x : w) x = ~0ULL; trim(); return *this; }
    dynamic_bitset &set(size_type i, bool v = true) { if (v) w[i >> 6] |= 1ULL << (i & 63); else w[i >> 6] &= ~(1ULL << (i & 63)); return *this; }
    dynamic_bitset &reset() { for (auto &x : w) x = 0; return *this; }
    dynamic_bitset &reset(size_type i) { return set(i, false); }
    dynamic_bitset &flip() { for (auto &x : w) x = ~x; trim(); return *this; }
    dynamic_bitset &flip(size_type i) { w[i >> 6] ^= 1ULL << (i & 63); return *this; }
    bool test(size_type i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    bool operator[](size_type i) const { return test(i); }
    class reference { // proxy of a single bit, as in boost (b[i] = x, b[i] |= x, ~b[i], bool(b[i]))
    public:
        reference(dynamic_bitset &b, size_type i) : b(b), i(i) {}
        reference &operator=(bool v) { b.set(i, v); return *this; }
        reference &operator=(const reference &o) { b.set(i, (bool)o); return *this; }
        reference &operator|=(bool v) { if (v) b.set(i); return *this; }
        reference &operator&=(bool v) { if (!v) b.reset(i); return *this; }
        reference &flip() { b.flip(i); return *this; }
        operator bool() const { return b.test(i); }
        bool operator~() const { return !b.test(i); }
    private:
        dynamic_bitset &b;
        size_type i;
    };
    reference operator[](size_type i) { return reference(*this, i); }
    bool any() const { for (auto x : w) if (x) return true; return false; }
    bool none() const { return !any(); }
    size_type count() const { size_type c = 0; for (auto x : w) c += (size_type)__builtin_popcountll(x); return c; }
    size_type find_first() const { for (size_type i = 0; i < w.size(); i++) if (w[i]) return i * 64 + (size_type)__builtin_ctzll(w[i]); return npos; }
    size_type find_next(size_type pos) const {
        pos++;
        if (pos >= nbits) return npos;
        size_type i = pos >> 6;
        uint64_t x = w[i] & (~0ULL << (pos & 63));
        while (true) {
            if (x) return i * 64 + (size_type)__builtin_ctzll(x);
            if (++i >= w.size()) return npos;
            x = w[i];
        }
    }
    dynamic_bitset &operator&=(const dynamic_bitset &o) { for (size_type i = 0; i < w.size(); i++) w[i] &= o.w[i]; return *this; }
    dynamic_bitset &operator|=(const dynamic_bitset &o) { for (size_type i = 0; i < w.size(); i++) w[i] |= o.w[i]; return *this; }
    dynamic_bitset &operator^=(const dynamic_bitset &o) { for (size_type i = 0; i < w.size(); i++) w[i] ^= o.w[i]; return *this; }
    dynamic_bitset &operator-=(const dynamic_bitset &o) { for (size_type i = 0; i < w.size(); i++) w[i] &= ~o.w[i]; return *this; }
    dynamic_bitset operator~() const { dynamic_bitset r(*this); r.flip(); return r; }
    bool operator==(const dynamic_bitset &o) const { return nbits == o.nbits && w == o.w; }
    bool operator!=(const dynamic_bitset &o) const { return !(*this == o); }
    bool operator<(const dynamic_bitset &o) const {
        for (size_type i = w.size(); i-- > 0;) if (w[i] != o.w[i]) return w[i] < o.w[i];
        return false;
    }
    bool is_subset_of(const dynamic_bitset &o) const { for (size_type i = 0; i < w.size(); i++) if (w[i] & ~o.w[i]) return false; return true; }
    unsigned long to_ulong() const { return w.empty() ? 0UL : (unsigned long)w[0]; }
private:
    void trim() { if (nbits & 63) w.back() &= (1ULL << (nbits & 63)) - 1; }
    std::vector<uint64_t> w;
    size_type nbits;
};
template <class B, class A> inline dynamic_bitset<B, A> operator&(const dynamic_bitset<B, A> &a, const dynamic_bitset<B, A> &b) { dynamic_bitset<B, A> r(a); r &= b; return r; }
template <class B, class A> inline dynamic_bitset<B, A> operator|(const dynamic_bitset<B, A> &a, const dynamic_bitset<B, A> &b) { dynamic_bitset<B, A> r(a); r |= b; return r; }
} // namespace boost
