// sdsl_b200.hpp — SDSL-shaped C++ classes over the C ABI (include/sdslgpu.h).
//
// Header-only host-side mirror of the reference's interface for the hot path: the same
// names, argument meaning and in-band results as xxsds/sdsl-lite, so that a call site written against
//     sdsl::bit_vector / rank_support_v<b> / select_support_mcl<b> / rrr_vector<63> / sd_vector<> /
//     wt_huff<> / wt_int<> / csa_wt<wt_huff<>> / sdsl::count / sdsl::locate
// compiles against namespace sdsl_b200 unchanged, and gains batch overloads (pointer + count, or
// std::vector) that are the fast path: one call = one kernel launch over the whole batch.
// Scalar calls (rank(i), wt.rank(i,c), count(csa, b, e)) are batches of one and cost a launch each: they
// exist for drop-in compatibility and tests, not for speed.
//
// Ownership mirrors the reference: supports hold a NON-OWNING pointer to their vector and are re-pointed
// with set_vector (rank_support.hpp:33, util.hpp:414-432).  Errors from the C ABI become std::runtime_error
// (the reference throws std::logic_error / std::runtime_error from constructors too).
#pragma once
#include <array>
#include <cstdint>
#include <fstream>
#include <istream>
#include <iterator>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/sdslgpu.h"

namespace sdsl_b200
{

typedef uint64_t size_type;

inline void check(int status, char const * what)
{
    if (status != SDSLGPU_OK)
        throw std::runtime_error(std::string(what) + ": " + sdslgpu_last_error());
}

namespace detail
{
struct handle_deleter
{
    void operator()(sdslgpu_handle * h) const
    {
        sdslgpu_free(h);
    }
};
typedef std::shared_ptr<sdslgpu_handle> handle_ptr;
inline handle_ptr adopt(sdslgpu_handle * h)
{
    return handle_ptr(h, handle_deleter());
}
inline int & default_device()
{
    static int d = 0;
    return d;
}
// the reference's serialize() bytes of a device image (sdslgpu_serialize) / a device image from such bytes
inline std::vector<uint8_t> image_blob(sdslgpu_handle const * h, int what)
{
    uint64_t n = 0;
    check(sdslgpu_serialize(h, what, nullptr, 0, &n), "serialize");
    std::vector<uint8_t> blob(n ? n : 1);
    check(sdslgpu_serialize(h, what, blob.data(), n, &n), "serialize");
    blob.resize(n);
    return blob;
}
inline size_type write_blob(sdslgpu_handle const * h, int what, std::ostream & out)
{
    std::vector<uint8_t> blob = image_blob(h, what);
    out.write(reinterpret_cast<char const *>(blob.data()), (std::streamsize)blob.size());
    return blob.size();
}
inline uint64_t read_u64(std::istream & in)
{
    uint64_t x = 0;
    in.read(reinterpret_cast<char *>(&x), 8);
    if (!in)
        throw std::runtime_error("load: truncated stream");
    return x;
}
// skips one serialised int_vector<w> (int_vector.hpp:884-916): header (width << 56 | bit_size), ceil(bit_size/64) words
inline void skip_int_vector(std::istream & in)
{
    uint64_t bits = read_u64(in) & ((1ull << 56) - 1);
    in.ignore((std::streamsize)(((bits + 63) >> 6) * 8));
    if (!in)
        throw std::runtime_error("load: truncated stream");
}
// Like the reference's load(std::istream&): consumes exactly the structure's bytes.  The library parses from a
// buffer, so the rest of the stream is read, sdslgpu_load_sdsl_ex reports how much of it the structure occupied, and
// the stream is put back right behind it (a stream that cannot seek keeps the old behaviour: read to its end).
inline handle_ptr load_blob(std::istream & in, int kind, uint32_t flags, uint32_t param, uint32_t isa_dens = 0)
{
    std::istream::pos_type const start = in.tellg();
    std::vector<char> blob((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>()); // the rest of the stream
    sdslgpu_handle * h = nullptr;
    uint64_t used = 0;
    check(sdslgpu_load_sdsl_ex(blob.data(), blob.size(), kind, default_device(), flags, param, isa_dens, &used, &h), "load");
    handle_ptr p = adopt(h);
    if (start != std::istream::pos_type(-1))
    {
        in.clear();
        in.seekg(start + (std::streamoff)used);
    }
    return p;
}
} // namespace detail

//! Device used by constructors that do not name one (there is no CPU fallback).
inline void set_device(int d)
{
    detail::default_device() = d;
}

// ------------------------------------------------------------------------------------------------------
// bit_vector = int_vector<1> (int_vector.hpp:210-257): host words + a lazily built device image that
// carries rank_support_v<0/1> and select_support_mcl<0/1>.
// ------------------------------------------------------------------------------------------------------
struct bv_tag;
template <uint8_t t_b, class t_vec, uint8_t t_pat_len = 1>
class rank_support;
template <uint8_t t_b, class t_vec, uint8_t t_pat_len = 1>
class select_support;
class bit_vector
{
public:
    typedef sdsl_b200::size_type size_type;
    typedef bv_tag index_category; // sdsl_concepts.hpp
    typedef bool value_type;
    typedef rank_support<1, bit_vector, 1> rank_1_type; // int_vector.hpp:227-231
    typedef rank_support<0, bit_vector, 1> rank_0_type;
    typedef select_support<1, bit_vector, 1> select_1_type;
    typedef select_support<0, bit_vector, 1> select_0_type;
    bit_vector() = default;
    explicit bit_vector(size_type n, bool value = false) : m_size(n), m_words((n + 63) / 64 + 1, value ? ~0ull : 0ull)
    {
        trim();
    }
    bit_vector(uint64_t const * words, size_type n) : m_size(n), m_words(words, words + (n + 63) / 64)
    {
        m_words.push_back(0);
    }
    size_type size() const
    {
        return m_size;
    }
    size_type bit_size() const
    {
        return m_size;
    }
    bool empty() const
    {
        return m_size == 0;
    }
    uint64_t const * data() const
    {
        return m_words.data();
    }
    uint64_t * data()
    {
        m_image.reset(); // the caller may write: the device image is rebuilt on next use
        return m_words.data();
    }
    bool operator[](size_type i) const // int_vector.hpp:1900-1904
    {
        return (m_words[i >> 6] >> (i & 63)) & 1;
    }
    //! bv[i] = b (int_vector_reference<bit_vector>, int_vector.hpp:1006-1070)
    class reference
    {
    public:
        reference(bit_vector * v, size_type i) : m_v(v), m_i(i)
        {}
        reference & operator=(bool b)
        {
            m_v->set(m_i, b);
            return *this;
        }
        reference & operator=(reference const & o)
        {
            return *this = (bool)o;
        }
        operator bool() const
        {
            return (m_v->m_words[m_i >> 6] >> (m_i & 63)) & 1;
        }

    private:
        bit_vector * m_v;
        size_type m_i;
    };
    reference operator[](size_type i)
    {
        return reference(this, i);
    }
    void set(size_type i, bool v)
    {
        m_image.reset();
        if (v)
            m_words[i >> 6] |= 1ull << (i & 63);
        else
            m_words[i >> 6] &= ~(1ull << (i & 63));
    }
    //! the device image (built on first use)
    sdslgpu_handle const * image() const
    {
        if (!m_image)
        {
            sdslgpu_handle * h = nullptr;
            check(sdslgpu_bv_create(m_words.data(), m_size, detail::default_device(), m_flags, &h), "bit_vector");
            m_image = detail::adopt(h);
        }
        return m_image.get();
    }
    //! int_vector<1>::serialize (int_vector.hpp:1995-2004): header (1 << 56 | size), then the words
    size_type serialize(std::ostream & out) const
    {
        uint64_t header = (1ull << 56) | m_size, nw = (m_size + 63) >> 6;
        out.write(reinterpret_cast<char const *>(&header), 8);
        out.write(reinterpret_cast<char const *>(m_words.data()), (std::streamsize)(nw * 8));
        return 8 + nw * 8;
    }
    //! int_vector<1>::load (int_vector.hpp:2006-2022): consumes exactly one serialised bit_vector
    void load(std::istream & in)
    {
        uint64_t header = detail::read_u64(in);
        if ((header >> 56) != 1)
            throw std::runtime_error("bit_vector::load: not an int_vector<1>");
        m_size = header & ((1ull << 56) - 1);
        m_words.assign(((m_size + 63) >> 6) + 1, 0);
        in.read(reinterpret_cast<char *>(m_words.data()), (std::streamsize)(((m_size + 63) >> 6) * 8));
        if (!in)
            throw std::runtime_error("bit_vector::load: truncated stream");
        m_image.reset();
    }
    bool operator==(bit_vector const & o) const
    {
        if (m_size != o.m_size)
            return false;
        for (size_type w = 0; w < (m_size >> 6); ++w)
            if (m_words[w] != o.m_words[w])
                return false;
        return (m_size & 63) == 0 || ((m_words[m_size >> 6] ^ o.m_words[m_size >> 6]) & ((1ull << (m_size & 63)) - 1)) == 0;
    }
    bool operator!=(bit_vector const & o) const
    {
        return !(*this == o);
    }
    //! Keep select on its samples instead of letting the first large select batch build select sectors (32 bytes per ~81
    //! ones of a half-dense vector; select in one gather, 1.25x faster): SDSLGPU_F_COMPACT, see sdslgpu_select in sdslgpu.h.
    //! Takes effect when the device image is (re)built; never changes a result.
    void compact_select(bool on = true)
    {
        uint32_t const f = on ? SDSLGPU_F_COMPACT : SDSLGPU_F_DEFAULT;
        if (f != m_flags)
        {
            m_flags = f;
            m_image.reset();
        }
    }
    //! how large batches are executed: SDSLGPU_ORDER_AUTO (default) / _DIRECT / _BINNED; never changes a result
    void batch_order(int order) const
    {
        check(sdslgpu_set_batch_order(const_cast<sdslgpu_handle *>(image()), order), "bit_vector::batch_order");
    }

private:
    void trim()
    {
        if (m_size & 63)
            m_words[m_size >> 6] &= (1ull << (m_size & 63)) - 1;
        m_words.back() = 0;
        if ((m_size & 63) == 0 && !m_words.empty())
            m_words[(m_size + 63) / 64] = 0;
    }
    size_type m_size = 0;
    std::vector<uint64_t> m_words;
    uint32_t m_flags = SDSLGPU_F_DEFAULT;
    mutable detail::handle_ptr m_image;
};

namespace detail
{
// shared implementation of the rank / select support concept (rank_support.hpp:30-75, select_support.hpp:32-78)
template <class t_vec>
class support_base
{
public:
    typedef sdsl_b200::size_type size_type;
    typedef t_vec bit_vector_type;
    explicit support_base(t_vec const * v = nullptr) : m_v(v)
    {}
    void set_vector(t_vec const * v = nullptr)
    {
        m_v = v;
    }
    size_type size() const
    {
        return m_v ? m_v->size() : 0;
    }
    //! supports are views of their vector's device image: two supports are equal when they answer for the same vector
    bool operator==(support_base const & o) const noexcept
    {
        return m_v == o.m_v;
    }
    bool operator!=(support_base const & o) const noexcept
    {
        return m_v != o.m_v;
    }

protected:
    sdslgpu_handle const * image() const
    {
        if (!m_v)
            throw std::runtime_error("support used without a vector (set_vector)");
        return m_v->image();
    }
    t_vec const * m_v;
};

//! <t_b, t_pat_len> of the reference (rank_support.hpp:105-284: 0, 1, and 10 / 01 / 00 / 11 with length 2; note that
//! the literals 01 and 00 are the integers 1 and 0) -> the C ABI's pattern code
constexpr int pattern_code(uint8_t t_b, uint8_t t_pat_len)
{
    return t_pat_len == 1 ? (t_b ? SDSLGPU_PAT_1 : SDSLGPU_PAT_0)
                          : (t_b == 10 ? SDSLGPU_PAT_10 : t_b == 1 ? SDSLGPU_PAT_01 : t_b == 0 ? SDSLGPU_PAT_00 : SDSLGPU_PAT_11);
}
} // namespace detail

//! rank_support_v<t_b,1> concept (rank_support_v.hpp:47-192) for any bit-vector type of this header.
template <uint8_t t_b, class t_vec, uint8_t t_pat_len>
class rank_support : public detail::support_base<t_vec>
{
    using base = detail::support_base<t_vec>;

public:
    using typename base::size_type;
    enum
    {
        bit_pat = t_b,
        bit_pat_len = t_pat_len
    };
    using base::base;
    size_type rank(size_type i) const
    {
        size_type r;
        rank(&i, 1, &r);
        return r;
    }
    size_type operator()(size_type i) const
    {
        return rank(i);
    }
    //! batch: out[k] = rank(idx[k]); pointers may be host or device memory
    void rank(uint64_t const * idx, size_type n, uint64_t * out, void * stream = nullptr) const
    {
        check(sdslgpu_rank(this->image(), detail::pattern_code(t_b, t_pat_len), idx, n, out, stream), "rank");
    }
    std::vector<uint64_t> rank(std::vector<uint64_t> const & idx) const
    {
        std::vector<uint64_t> out(idx.size());
        rank(idx.data(), idx.size(), out.data());
        return out;
    }
    //! serialize: over a plain bit_vector the reference's m_basic_block, byte for byte (rank_support_v.hpp:151-158);
    //! the supports of rrr_vector / sd_vector serialise to nothing, as in the reference (rrr_vector.hpp:580-585)
    size_type serialize(std::ostream & out) const
    {
        if (!std::is_same<t_vec, bit_vector>::value)
            return 0;
        if (t_pat_len != 1)
            throw std::runtime_error("rank_support::serialize: only the one-bit patterns have a serialised form here");
        return detail::write_blob(this->image(), t_b ? 1 : 2, out);
    }
    //! load(in, v) (rank_support_v.hpp:160-165): the stored table is skipped — the device image of *v answers
    void load(std::istream & in, t_vec const * v = nullptr)
    {
        this->set_vector(v);
        if (std::is_same<t_vec, bit_vector>::value)
            detail::skip_int_vector(in);
    }
};

template <uint8_t t_b, class t_vec, uint8_t t_pat_len>
class select_support : public detail::support_base<t_vec>
{
    using base = detail::support_base<t_vec>;

public:
    using typename base::size_type;
    enum
    {
        bit_pat = t_b,
        bit_pat_len = t_pat_len
    };
    using base::base;
    size_type select(size_type i) const
    {
        size_type r;
        select(&i, 1, &r);
        return r;
    }
    size_type operator()(size_type i) const
    {
        return select(i);
    }
    void select(uint64_t const * i, size_type n, uint64_t * out, void * stream = nullptr) const
    {
        check(sdslgpu_select(this->image(), detail::pattern_code(t_b, t_pat_len), i, n, out, stream), "select");
    }
    std::vector<uint64_t> select(std::vector<uint64_t> const & i) const
    {
        std::vector<uint64_t> out(i.size());
        select(i.data(), i.size(), out.data());
        return out;
    }
    //! serialize: over a plain bit_vector the reference's select_support_mcl bytes (select_support_mcl.hpp:474-518)
    size_type serialize(std::ostream & out) const
    {
        if (!std::is_same<t_vec, bit_vector>::value)
            return 0;
        if (t_pat_len != 1)
            throw std::runtime_error("select_support::serialize: only the one-bit patterns have a serialised form here");
        return detail::write_blob(this->image(), t_b ? 3 : 4, out);
    }
    //! load(in, v) (select_support_mcl.hpp:521-555): the stored samples are skipped — the device image of *v answers
    void load(std::istream & in, t_vec const * v = nullptr)
    {
        this->set_vector(v);
        if (!std::is_same<t_vec, bit_vector>::value)
            return;
        uint64_t arg_cnt = detail::read_u64(in);
        if (arg_cnt == 0)
            return;
        detail::skip_int_vector(in); // m_superblock
        detail::skip_int_vector(in); // mini_or_long
        for (uint64_t sb = (arg_cnt + 4095) >> 12; sb > 0; --sb)
            detail::skip_int_vector(in); // one mini or long block per superblock
    }
};

template <uint8_t t_b = 1, uint8_t t_pat_len = 1>
using rank_support_v = rank_support<t_b, bit_vector, t_pat_len>;
template <uint8_t t_b = 1, uint8_t t_pat_len = 1>
using select_support_mcl = select_support<t_b, bit_vector, t_pat_len>;
//! rank_support_v5 (rank_support_v5.hpp:131-149) answers exactly what rank_support_v answers; on the device both
//! map onto the same sector blocks.  Only the serialised form differs: the 6.25 %-overhead table of 2048-bit
//! superblocks (rank_support_v5.hpp:66-122), rebuilt on the device, byte for byte.
template <uint8_t t_b = 1, uint8_t t_pat_len = 1>
class rank_support_v5 : public rank_support<t_b, bit_vector, t_pat_len>
{
public:
    using rank_support<t_b, bit_vector, t_pat_len>::rank_support;
    size_type serialize(std::ostream & out) const
    {
        if (t_pat_len != 1)
            throw std::runtime_error("rank_support_v5::serialize: only the one-bit patterns have a serialised form here");
        return detail::write_blob(this->image(), t_b ? 5 : 6, out);
    }
};

// ------------------------------------------------------------------------------------------------------
// compressed bit vectors (rrr_vector.hpp:67-109, sd_vector.hpp:131-163)
// ------------------------------------------------------------------------------------------------------
namespace detail
{
template <int KIND>
class compressed_vector
{
public:
    typedef sdsl_b200::size_type size_type;
    typedef rank_support<1, compressed_vector> rank_1_type;
    typedef rank_support<0, compressed_vector> rank_0_type;
    typedef select_support<1, compressed_vector> select_1_type;
    typedef select_support<0, compressed_vector> select_0_type;
    typedef bv_tag index_category;
    typedef bool value_type;
    compressed_vector() = default;
    explicit compressed_vector(bit_vector const & bv) : m_size(bv.size())
    {
        sdslgpu_handle * h = nullptr;
        if (KIND == SDSLGPU_KIND_RRR63)
            check(sdslgpu_rrr63_create(bv.data(), bv.size(), default_device(), SDSLGPU_F_DEFAULT, &h), "rrr_vector");
        else
            check(sdslgpu_sd_create(bv.data(), bv.size(), default_device(), SDSLGPU_F_DEFAULT, &h), "sd_vector");
        m_image = adopt(h);
    }
    size_type size() const
    {
        return m_size;
    }
    bool operator[](size_type i) const
    {
        uint64_t r;
        check(sdslgpu_access(image(), &i, 1, &r, nullptr), "operator[]");
        return r != 0;
    }
    //! serialize (rrr_vector.hpp:366-378, sd_vector.hpp:426-438): the reference's bytes, supports included
    size_type serialize(std::ostream & out) const
    {
        return detail::write_blob(image(), KIND == SDSLGPU_KIND_SD ? 1 : 0, out);
    }
    //! load: reads the REST of the stream (one structure per stream / file, as store_to_file writes it)
    void load(std::istream & in)
    {
        m_image = detail::load_blob(in, KIND, SDSLGPU_F_DEFAULT, 0);
        check(sdslgpu_size(m_image.get(), &m_size), "size");
    }
    sdslgpu_handle const * image() const
    {
        if (!m_image)
            throw std::runtime_error("empty compressed vector");
        return m_image.get();
    }

private:
    size_type m_size = 0;
    handle_ptr m_image;
};
} // namespace detail

//! rrr_vector<t_bs, t_rac, t_k> (rrr_vector.hpp:67-76): only t_bs = 63, t_k = 32 is on the hot path
template <uint16_t t_bs = 63, class t_rac = void, uint16_t t_k = 32>
class rrr_vector : public detail::compressed_vector<SDSLGPU_KIND_RRR63>
{
    static_assert(t_bs == 63 && t_k == 32, "this engine implements rrr_vector<63, int_vector<>, 32> (the benchmarked block size)");

public:
    using detail::compressed_vector<SDSLGPU_KIND_RRR63>::compressed_vector;
};
//! sd_vector<t_hi_bit_vector, t_select_1, t_select_0> (sd_vector.hpp:131-136): the template arguments choose host-side
//! component types in the reference; on the device `high` always is a sector-block image
template <class t_hi_bit_vector = bit_vector, class t_select_1 = void, class t_select_0 = void>
class sd_vector : public detail::compressed_vector<SDSLGPU_KIND_SD>
{
public:
    using detail::compressed_vector<SDSLGPU_KIND_SD>::compressed_vector;
};
//! select_0_support_sd (sd_vector.hpp:752-978): select_0 straight from sampled pointers into `high`; here every
//! select_support<0, sd_vector> already answers that way (csrc/sd_device.cuh sd_select0_one)
template <class t_sd_vector = sd_vector<>>
using select_0_support_sd = select_support<0, detail::compressed_vector<SDSLGPU_KIND_SD>>;

// ------------------------------------------------------------------------------------------------------
// category tags (sdsl_concepts.hpp:16-44) and the random-access iterator the containers hand out (iterators.hpp:21-140)
// ------------------------------------------------------------------------------------------------------
struct bv_tag {};
struct iv_tag {};
struct csa_tag {};
struct cst_tag {};
struct wt_tag {};
struct psi_tag {};
struct lf_tag {};
struct csa_member_tag {};
struct byte_alphabet_tag
{
    static constexpr uint8_t WIDTH = 8;
};
struct int_alphabet_tag
{
    static constexpr uint8_t WIDTH = 0;
};

template <class t_container>
class random_access_const_iterator
{
public:
    typedef std::random_access_iterator_tag iterator_category;
    typedef decltype(std::declval<t_container const &>()[0]) value_type;
    typedef std::ptrdiff_t difference_type;
    typedef void pointer;
    typedef value_type reference;
    random_access_const_iterator() = default;
    random_access_const_iterator(t_container const * c, size_type i = 0) : m_c(c), m_i(i)
    {}
    value_type operator*() const
    {
        return (*m_c)[m_i];
    }
    value_type operator[](difference_type d) const
    {
        return (*m_c)[m_i + d];
    }
    random_access_const_iterator & operator++()
    {
        ++m_i;
        return *this;
    }
    random_access_const_iterator operator++(int)
    {
        random_access_const_iterator t = *this;
        ++m_i;
        return t;
    }
    random_access_const_iterator & operator--()
    {
        --m_i;
        return *this;
    }
    random_access_const_iterator & operator+=(difference_type d)
    {
        m_i += d;
        return *this;
    }
    random_access_const_iterator & operator-=(difference_type d)
    {
        m_i -= d;
        return *this;
    }
    random_access_const_iterator operator+(difference_type d) const
    {
        return random_access_const_iterator(m_c, m_i + d);
    }
    random_access_const_iterator operator-(difference_type d) const
    {
        return random_access_const_iterator(m_c, m_i - d);
    }
    difference_type operator-(random_access_const_iterator const & o) const
    {
        return (difference_type)m_i - (difference_type)o.m_i;
    }
    bool operator==(random_access_const_iterator const & o) const
    {
        return m_i == o.m_i;
    }
    bool operator!=(random_access_const_iterator const & o) const
    {
        return m_i != o.m_i;
    }
    bool operator<(random_access_const_iterator const & o) const
    {
        return m_i < o.m_i;
    }

private:
    t_container const * m_c = nullptr;
    size_type m_i = 0;
};

// ------------------------------------------------------------------------------------------------------
// wavelet trees (wt_pc.hpp:61-78, 314-474; wt_int.hpp:57-61, 340-507)
// ------------------------------------------------------------------------------------------------------
namespace detail
{
template <class t_sym>
class wavelet_tree_base
{
public:
    typedef sdsl_b200::size_type size_type;
    typedef t_sym value_type;
    typedef wt_tag index_category;
    typedef random_access_const_iterator<wavelet_tree_base> const_iterator;
    typedef const_iterator iterator;
    size_type size() const
    {
        return m_size;
    }
    bool empty() const
    {
        return m_size == 0;
    }
    size_type sigma = 0;
    value_type operator[](size_type i) const
    {
        uint64_t s;
        check(sdslgpu_wt_access(image(), &i, 1, &s, nullptr, nullptr), "wt[i]");
        return (value_type)s;
    }
    const_iterator begin() const // wt_pc.hpp:694-701: the symbols in order (one scalar call each: for compatibility)
    {
        return const_iterator(this, 0);
    }
    const_iterator end() const
    {
        return const_iterator(this, m_size);
    }
    size_type rank(size_type i, value_type c) const
    {
        size_type r;
        rank(&i, &c, 1, &r);
        return r;
    }
    size_type select(size_type i, value_type c) const
    {
        size_type r;
        select(&i, &c, 1, &r);
        return r;
    }
    std::pair<size_type, value_type> inverse_select(size_type i) const
    {
        uint64_t s, r;
        check(sdslgpu_wt_access(image(), &i, 1, &s, &r, nullptr), "inverse_select");
        return std::make_pair((size_type)r, (value_type)s);
    }
    // batch forms
    void rank(uint64_t const * i, value_type const * c, size_type n, uint64_t * out, void * stream = nullptr) const
    {
        check(sdslgpu_wt_rank(image(), i, c, n, out, stream), "wt.rank");
    }
    void select(uint64_t const * i, value_type const * c, size_type n, uint64_t * out, void * stream = nullptr) const
    {
        check(sdslgpu_wt_select(image(), i, c, n, out, stream), "wt.select");
    }
    void access(uint64_t const * i, size_type n, uint64_t * sym_out, uint64_t * rank_out = nullptr, void * stream = nullptr) const
    {
        check(sdslgpu_wt_access(image(), i, n, sym_out, rank_out, stream), "wt.access");
    }
    //! serialize (wt_pc.hpp:713-726, wt_int.hpp:792-805): byte-identical to the reference's, so that the reference
    //! can load an index built here (and `load` below takes what the reference stored)
    size_type serialize(std::ostream & out) const
    {
        return detail::write_blob(image(), 0, out);
    }
    sdslgpu_handle const * image() const
    {
        if (!m_image)
            throw std::runtime_error("empty wavelet tree");
        return m_image.get();
    }
    //! a view of an image somebody else owns (csa.wavelet_tree / csa.bwt share the index's handle)
    void share_image(handle_ptr p)
    {
        adopt_ptr(p);
    }

protected:
    void adopt_image(sdslgpu_handle * h)
    {
        adopt_ptr(adopt(h));
    }
    void adopt_ptr(handle_ptr p)
    {
        m_image = p;
        check(sdslgpu_size(p.get(), &m_size), "size");
        uint64_t s = 0;
        check(sdslgpu_wt_sigma(p.get(), &s), "sigma");
        sigma = s;
    }
    size_type m_size = 0;
    handle_ptr m_image;
};

template <class t_bitvector>
struct wt_flags
{
    static constexpr uint32_t value = SDSLGPU_F_DEFAULT;
};
template <uint16_t t_bs, class t_rac, uint16_t t_k>
struct wt_flags<rrr_vector<t_bs, t_rac, t_k>>
{
    static constexpr uint32_t value = SDSLGPU_F_RRR_BV; // wt_huff<rrr_vector<63>>: the H0-compressed tree
};
} // namespace detail

//! wt_huff<t_bitvector, t_rank, t_select, t_select_zero, t_tree_strat> (wt_huff.hpp:56-67).  t_bitvector = bit_vector
//! (default) or rrr_vector<63>; the support types are host-side choices of the reference — the device image carries
//! its own rank blocks and select samples whatever they are.
template <class t_bitvector = bit_vector, class t_rank = void, class t_select = void, class t_select_zero = void, class t_tree_strat = void>
class wt_huff : public detail::wavelet_tree_base<uint8_t>
{
public:
    typedef byte_alphabet_tag alphabet_category;
    typedef t_bitvector bit_vector_type;
    enum
    {
        lex_ordered = 0 // Huffman shape: not lexicographically ordered (wt_huff.hpp:49-53)
    };
    static constexpr uint32_t image_flags = detail::wt_flags<t_bitvector>::value;
    wt_huff() = default;
    //! wt_pc(t_it begin, t_it end, tmp_dir) (wt_pc.hpp:194): any range of bytes
    template <class t_it>
    wt_huff(t_it begin, t_it end, std::string const & = std::string())
    {
        std::vector<uint8_t> text(begin, end);
        sdslgpu_handle * h = nullptr;
        check(sdslgpu_wt_huff_create(text.data(), (uint64_t)text.size(), detail::default_device(), image_flags, &h), "wt_huff");
        adopt_image(h);
    }
    explicit wt_huff(std::string const & text) : wt_huff(text.begin(), text.end())
    {}
    //! load (wt_pc.hpp:729-741): consumes exactly the tree's bytes
    void load(std::istream & in)
    {
        adopt_ptr(detail::load_blob(in, SDSLGPU_KIND_WT_HUFF, image_flags, 0));
    }
};

template <class t_bitvector = bit_vector, class t_rank = void, class t_select = void, class t_select_zero = void>
class wt_int : public detail::wavelet_tree_base<uint64_t>
{
public:
    typedef int_alphabet_tag alphabet_category;
    enum
    {
        lex_ordered = 1
    };
    wt_int() = default;
    template <class t_it>
    wt_int(t_it begin, t_it end, std::string const & = std::string())
    {
        std::vector<uint64_t> seq(begin, end);
        sdslgpu_handle * h = nullptr;
        check(sdslgpu_wt_int_create(seq.data(), (uint64_t)seq.size(), detail::default_device(), SDSLGPU_F_DEFAULT, &h), "wt_int");
        adopt_image(h);
    }
    explicit wt_int(std::vector<uint64_t> const & seq) : wt_int(seq.begin(), seq.end())
    {}
    //! load (wt_int.hpp:808-821): consumes exactly the tree's bytes
    void load(std::istream & in)
    {
        adopt_ptr(detail::load_blob(in, SDSLGPU_KIND_WT_INT, SDSLGPU_F_DEFAULT, 0));
    }
};

// ------------------------------------------------------------------------------------------------------
// csa_wt<wt_huff<>> and the search algorithms (csa_wt.hpp:49-130; suffix_array_algorithm.hpp:166-248,463-570)
// ------------------------------------------------------------------------------------------------------
namespace detail
{
//! csa.bwt (suffix_array_helper.hpp:424-470): bwt[i], bwt.rank(i, c), bwt.select(i, c), size()
class bwt_view
{
public:
    typedef sdsl_b200::size_type size_type;
    typedef uint8_t value_type;
    typedef csa_member_tag category;
    typedef random_access_const_iterator<bwt_view> const_iterator;
    bwt_view() = default;
    explicit bwt_view(handle_ptr p, size_type n) : m_image(p), m_size(n)
    {}
    size_type size() const
    {
        return m_size;
    }
    bool empty() const
    {
        return m_size == 0;
    }
    value_type operator[](size_type i) const
    {
        uint64_t s;
        check(sdslgpu_wt_access(m_image.get(), &i, 1, &s, nullptr, nullptr), "bwt[i]");
        return (value_type)s;
    }
    size_type rank(size_type i, value_type c) const
    {
        uint64_t r;
        check(sdslgpu_wt_rank(m_image.get(), &i, &c, 1, &r, nullptr), "bwt.rank");
        return r;
    }
    size_type select(size_type i, value_type c) const
    {
        uint64_t r;
        check(sdslgpu_wt_select(m_image.get(), &i, &c, 1, &r, nullptr), "bwt.select");
        return r;
    }
    const_iterator begin() const
    {
        return const_iterator(this, 0);
    }
    const_iterator end() const
    {
        return const_iterator(this, m_size);
    }

private:
    handle_ptr m_image;
    size_type m_size = 0;
};

//! csa.lf (suffix_array_helper.hpp:346-360): lf[i] = C[char2comp[bwt[i]]] + bwt.rank(i, bwt[i])
class lf_view
{
public:
    typedef sdsl_b200::size_type size_type;
    typedef size_type value_type;
    typedef csa_member_tag category;
    lf_view() = default;
    lf_view(handle_ptr p, size_type n, std::vector<uint64_t> const * C, std::vector<uint8_t> const * c2c) : m_image(p), m_size(n), m_C(C), m_c2c(c2c)
    {}
    size_type size() const
    {
        return m_size;
    }
    value_type operator[](size_type i) const
    {
        uint64_t s, r;
        check(sdslgpu_wt_access(m_image.get(), &i, 1, &s, &r, nullptr), "lf[i]"); // inverse_select: (rank(i, bwt[i]), bwt[i])
        return (*m_C)[(*m_c2c)[s]] + r;
    }

private:
    handle_ptr m_image;
    size_type m_size = 0;
    std::vector<uint64_t> const * m_C = nullptr;
    std::vector<uint8_t> const * m_c2c = nullptr;
};
} // namespace detail

//! csa_wt<t_wt, t_dens, t_inv_dens, t_sa_sample_strat, t_isa_sample_strat, t_alphabet_strat> (csa_wt.hpp:49-56) over a
//! byte alphabet.  t_wt = wt_huff<> or wt_huff<rrr_vector<63>>; the sampling strategies are the reference's defaults
//! (sa_order_sa_sampling, isa_sampling).  The public members backward_search / count / locate are written against in
//! the reference — char2comp, comp2char, C, sigma, bwt, lf, wavelet_tree (csa_wt.hpp:117-130) — are here with the same
//! names and meaning.
template <class t_wt = wt_huff<>, uint32_t t_dens = 32, uint32_t t_inv_dens = 64, class t_sa_sample_strat = void, class t_isa_sample_strat = void,
          class t_alphabet_strat = void>
class csa_wt
{
public:
    enum
    {
        sa_sample_dens = t_dens,
        isa_sample_dens = t_inv_dens
    };
    typedef sdsl_b200::size_type size_type;
    typedef uint64_t value_type;
    typedef std::ptrdiff_t difference_type;
    typedef uint8_t char_type;
    typedef uint8_t comp_char_type;
    typedef std::string string_type;
    typedef t_wt wavelet_tree_type;
    typedef detail::bwt_view bwt_type;
    typedef detail::lf_view lf_type;
    typedef csa_tag index_category;
    typedef lf_tag extract_category;
    typedef byte_alphabet_tag alphabet_category;
    typedef random_access_const_iterator<csa_wt> const_iterator;
    typedef const_iterator iterator;

    std::vector<uint8_t> char2comp; // csa_wt.hpp:117-120: 256 / sigma / sigma + 1 entries
    std::vector<uint8_t> comp2char;
    std::vector<uint64_t> C;
    uint16_t sigma = 0;
    bwt_type bwt; // csa_wt.hpp:122-127
    bwt_type L;
    lf_type lf;
    wavelet_tree_type wavelet_tree;

    csa_wt() = default;
    //! the text must not contain a 0 byte (construct.hpp:34-46); what construct_im(csa, text, 1) does
    explicit csa_wt(std::string const & text)
    {
        sdslgpu_handle * h = nullptr;
        uint32_t const flags = t_wt::image_flags;
        check(sdslgpu_csa_create_ex(reinterpret_cast<uint8_t const *>(text.data()), text.size(), detail::default_device(), flags, t_dens, t_inv_dens, &h),
              "csa_wt");
        bind(detail::adopt(h));
    }
    csa_wt(csa_wt const & o)
    {
        *this = o;
    }
    csa_wt & operator=(csa_wt const & o)
    {
        if (this != &o)
        {
            if (o.m_image)
                bind(o.m_image); // views are re-pointed at this object's own tables (util::init_support in the reference)
            else
                clear();
        }
        return *this;
    }
    csa_wt(csa_wt && o) noexcept
    {
        if (o.m_image)
            bind(o.m_image);
    }
    csa_wt & operator=(csa_wt && o) noexcept
    {
        if (this != &o && o.m_image)
            bind(o.m_image);
        return *this;
    }
    size_type size() const
    {
        return m_size;
    }
    bool empty() const
    {
        return m_size == 0;
    }
    static size_type max_size()
    {
        return ~(size_type)0 >> 8;
    }
    //! csa[i]: the i-th suffix array entry (csa_wt.hpp:363-381)
    value_type operator[](size_type i) const
    {
        uint64_t r;
        check(sdslgpu_fm_sa(image(), &i, 1, &r, nullptr), "csa[i]");
        return r;
    }
    const_iterator begin() const
    {
        return const_iterator(this, 0);
    }
    const_iterator end() const
    {
        return const_iterator(this, m_size);
    }
    //! csa.bwt.rank(i, c) (suffix_array_helper.hpp:461-464); kept from the first version of this header
    size_type rank_bwt(size_type i, char_type c) const
    {
        return bwt.rank(i, c);
    }
    //! serialize (csa_wt.hpp:389-402): wavelet tree, SA samples, ISA samples, alphabet — the reference's bytes
    size_type serialize(std::ostream & out) const
    {
        return detail::write_blob(image(), 0, out);
    }
    //! load (csa_wt.hpp:410-416): consumes exactly the index's bytes.  The densities are template arguments, as in the
    //! reference, and are checked against the sample counts in the file.
    //! flags: SDSLGPU_F_V5_SCAN for the reference's wt_huff<bit_vector, rank_support_v5<>, select_support_scan<>, ...>
    //! form (its count benchmark's FM_HUFF), SDSLGPU_F_COMPACT to keep the tree as the only occurrence structure
    void load(std::istream & in, uint32_t flags = SDSLGPU_F_DEFAULT)
    {
        bind(detail::load_blob(in, SDSLGPU_KIND_CSA_WT, flags | t_wt::image_flags, t_dens, t_inv_dens));
    }
    sdslgpu_handle const * image() const
    {
        if (!m_image)
            throw std::runtime_error("empty csa");
        return m_image.get();
    }

private:
    void clear()
    {
        m_image.reset();
        m_size = 0;
        sigma = 0;
        char2comp.clear();
        comp2char.clear();
        C.clear();
        bwt = L = bwt_type();
        lf = lf_type();
        wavelet_tree = wavelet_tree_type();
    }
    void bind(detail::handle_ptr p)
    {
        m_image = p;
        check(sdslgpu_size(p.get(), &m_size), "size");
        uint64_t c[257];
        uint8_t c2c[256], cc2[256];
        uint32_t sg = 0;
        check(sdslgpu_csa_alphabet(p.get(), c, c2c, cc2, &sg), "alphabet");
        sigma = (uint16_t)sg;
        char2comp.assign(c2c, c2c + 256);
        comp2char.assign(cc2, cc2 + sg);
        C.assign(c, c + sg + 1);
        bwt = bwt_type(p, m_size);
        L = bwt;
        lf = lf_type(p, m_size, &C, &char2comp);
        wavelet_tree.share_image(p);
    }
    size_type m_size = 0;
    detail::handle_ptr m_image;
};

//! construct_im(idx, data, num_bytes) (construct.hpp:69-92) and construct(idx, file, num_bytes) (construct.hpp:127-193)
//! for the index types of this header: num_bytes must be 1 (a byte sequence).  Construction itself runs on the
//! device (suffix array by prefix doubling, BWT, samples, wavelet tree).
template <class t_wt, uint32_t t_dens, uint32_t t_inv_dens, class A, class B, class Cc>
void construct_im(csa_wt<t_wt, t_dens, t_inv_dens, A, B, Cc> & idx, std::string const & data, uint8_t num_bytes = 1)
{
    if (num_bytes != 1)
        throw std::runtime_error("construct_im: byte alphabets only (num_bytes = 1)");
    idx = csa_wt<t_wt, t_dens, t_inv_dens, A, B, Cc>(data);
}
template <class t_bv, class A, class B, class Cc, class D>
void construct_im(wt_huff<t_bv, A, B, Cc, D> & idx, std::string const & data, uint8_t num_bytes = 1)
{
    if (num_bytes != 1)
        throw std::runtime_error("construct_im: byte alphabets only (num_bytes = 1)");
    idx = wt_huff<t_bv, A, B, Cc, D>(data.begin(), data.end());
}
template <class t_index>
void construct(t_index & idx, std::string const & file, uint8_t num_bytes = 1)
{
    std::ifstream in(file, std::ios::binary);
    if (!in)
        throw std::runtime_error("construct: cannot open " + file);
    std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    construct_im(idx, data, num_bytes);
}

//! sdsl::count(csa, begin, end) (suffix_array_algorithm.hpp:463-471)
template <class t_csa, class t_pat_iter, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
typename t_csa::size_type count(t_csa const & csa, t_pat_iter begin, t_pat_iter end)
{
    std::string p(begin, end);
    uint64_t off[2] = {0, p.size()}, cnt = 0;
    check(sdslgpu_fm_count(csa.image(), reinterpret_cast<uint8_t const *>(p.data()), off, 1, &cnt, nullptr, nullptr), "count");
    return cnt;
}
template <class t_csa, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
typename t_csa::size_type count(t_csa const & csa, typename t_csa::string_type const & pat)
{
    return count(csa, pat.begin(), pat.end());
}

//! backward_search(csa, l, r, c, l_res, r_res) (suffix_array_algorithm.hpp:166-201): one step, the two bwt.rank calls of
//! the reference in one launch.  Returns the size of the new interval.
template <class t_csa, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
typename t_csa::size_type backward_search(t_csa const & csa, typename t_csa::size_type l, typename t_csa::size_type r, typename t_csa::char_type c,
                                          typename t_csa::size_type & l_res, typename t_csa::size_type & r_res)
{
    typename t_csa::size_type const cc = csa.char2comp[c];
    if (cc == 0 && c > 0)
    { // character is not in the text: the empty interval [1, 0]
        l_res = 1;
        r_res = 0;
        return 0;
    }
    typename t_csa::size_type const c_begin = csa.C[cc];
    if (l == 0 && r + 1 == csa.size())
    { // the whole suffix array: the table alone answers
        l_res = c_begin;
        r_res = csa.C[cc + 1] - 1;
    }
    else
    {
        uint64_t pos[2] = {l, r + 1}, rk[2];
        uint8_t sym[2] = {c, c};
        check(sdslgpu_wt_rank(csa.image(), pos, sym, 2, rk, nullptr), "backward_search");
        l_res = c_begin + rk[0];
        r_res = c_begin + rk[1] - 1;
    }
    return r_res + 1 - l_res;
}

//! backward_search(csa, l, r, begin, end, l_res, r_res) (suffix_array_algorithm.hpp:227-248): a whole pattern from an
//! arbitrary interval; from the full interval it is one call of the count kernel, otherwise step by step
template <class t_csa, class t_pat_iter, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
typename t_csa::size_type backward_search(t_csa const & csa, typename t_csa::size_type l, typename t_csa::size_type r, t_pat_iter begin, t_pat_iter end,
                                          typename t_csa::size_type & l_res, typename t_csa::size_type & r_res)
{
    t_pat_iter it = end;
    while (begin < it && r + 1 - l > 0)
    {
        --it;
        backward_search(csa, l, r, (typename t_csa::char_type) * it, l_res, r_res);
        l = l_res;
        r = r_res;
    }
    l_res = l;
    r_res = r;
    return r + 1 - l;
}

//! lex_interval(csa, begin, end) (suffix_array_algorithm.hpp:512-517): {l, r} of the pattern, r < l if it does not occur
template <class t_csa, class t_pat_iter, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
std::array<typename t_csa::size_type, 2> lex_interval(t_csa const & csa, t_pat_iter begin, t_pat_iter end)
{
    std::array<typename t_csa::size_type, 2> res;
    backward_search(csa, 0, csa.size() - 1, begin, end, res[0], res[1]);
    return res;
}

//! sdsl::locate(csa, begin, end): all occurrences, in suffix-array order (suffix_array_algorithm.hpp:534-550)
template <class t_csa, class t_pat_iter, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
std::vector<uint64_t> locate(t_csa const & csa, t_pat_iter begin, t_pat_iter end)
{
    std::string p(begin, end);
    uint64_t off[2] = {0, p.size()}, occ_off[2], total = 0;
    uint8_t const * bytes = reinterpret_cast<uint8_t const *>(p.data());
    check(sdslgpu_fm_locate(csa.image(), bytes, off, 1, occ_off, nullptr, 0, &total, nullptr), "locate");
    std::vector<uint64_t> occ(total);
    if (total)
        check(sdslgpu_fm_locate(csa.image(), bytes, off, 1, occ_off, occ.data(), total, &total, nullptr), "locate");
    return occ;
}
template <class t_csa, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
std::vector<uint64_t> locate(t_csa const & csa, typename t_csa::string_type const & pat)
{
    return locate(csa, pat.begin(), pat.end());
}

//! sdsl::extract(csa, begin, end): text[begin..end], end inclusive (suffix_array_algorithm.hpp:645-665)
template <class t_csa, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
typename t_csa::string_type extract(t_csa const & csa, typename t_csa::size_type begin, typename t_csa::size_type end)
{
    std::string out(end - begin + 1, '\0');
    uint64_t off[2] = {0, end - begin + 1};
    check(sdslgpu_fm_extract(csa.image(), &begin, &end, 1, off, reinterpret_cast<uint8_t *>(&out[0]), nullptr), "extract");
    return out;
}

//! batch count: one launch for the whole pattern set
template <class t_csa, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
std::vector<uint64_t> count(t_csa const & csa, std::vector<std::string> const & pats)
{
    std::string flat;
    std::vector<uint64_t> off(pats.size() + 1, 0), cnt(pats.size());
    for (size_t k = 0; k < pats.size(); ++k)
    {
        flat += pats[k];
        off[k + 1] = flat.size();
    }
    if (!pats.empty())
        check(sdslgpu_fm_count(csa.image(), reinterpret_cast<uint8_t const *>(flat.data()), off.data(), pats.size(), cnt.data(), nullptr, nullptr), "count");
    return cnt;
}

//! batch locate: occurrences of pattern k are occ[occ_off[k] .. occ_off[k+1]) in suffix-array order
template <class t_csa, class = typename std::enable_if<std::is_same<typename t_csa::index_category, csa_tag>::value>::type>
void locate(t_csa const & csa, std::vector<std::string> const & pats, std::vector<uint64_t> & occ_off, std::vector<uint64_t> & occ)
{
    std::string flat;
    std::vector<uint64_t> off(pats.size() + 1, 0);
    for (size_t k = 0; k < pats.size(); ++k)
    {
        flat += pats[k];
        off[k + 1] = flat.size();
    }
    occ_off.assign(pats.size() + 1, 0);
    uint64_t total = 0;
    uint8_t const * bytes = reinterpret_cast<uint8_t const *>(flat.data());
    check(sdslgpu_fm_locate(csa.image(), bytes, off.data(), pats.size(), occ_off.data(), nullptr, 0, &total, nullptr), "locate");
    occ.assign(total, 0);
    if (total)
        check(sdslgpu_fm_locate(csa.image(), bytes, off.data(), pats.size(), occ_off.data(), occ.data(), total, &total, nullptr), "locate");
}

//! sdsl::store_to_file(v, file) (io.hpp:877-896): true on success
template <class T>
bool store_to_file(T const & v, std::string const & file)
{
    std::ofstream out(file, std::ios::binary | std::ios::trunc | std::ios::out);
    if (!out)
        return false;
    v.serialize(out);
    out.close();
    return (bool)out;
}

//! sdsl::load_from_file(v, file) (io.hpp:992-1011): true on success
template <class T>
bool load_from_file(T & v, std::string const & file)
{
    std::ifstream in(file, std::ios::binary | std::ios::in);
    if (!in)
        return false;
    v.load(in);
    return true;
}

//! sdsl::size_in_bytes(v) (io.hpp:778-786): length of the serialised form
template <class T>
size_type size_in_bytes(T const & v)
{
    std::ostringstream os;
    return v.serialize(os);
}
//! sdsl::size_in_mega_bytes(v) (io.hpp:788-794)
template <class T>
double size_in_mega_bytes(T const & v)
{
    return (double)size_in_bytes(v) / (1024.0 * 1024.0);
}

} // namespace sdsl_b200
