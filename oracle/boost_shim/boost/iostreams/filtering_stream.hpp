// Minimal stand-in for the Boost.Iostreams pieces the reference touches (src/file_utils.{hpp,cpp}):
// filtering_istream + gzip_decompressor over a std::istream, filtering_ostream + gzip_compressor over
// a file_sink, and inert filtering_streambuf/copy/close for the reference's dead [[deprecated]] helpers.
// zlib does the work. Oracle build only - never linked into the product.
#pragma once
#include <zlib.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <ios>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <streambuf>
#include <string>
#include <vector>

namespace boost { namespace iostreams {

struct input {};
struct output {};
struct gzip_decompressor {};
struct gzip_compressor {};
struct file_sink {
    std::string name;
    std::ios_base::openmode mode;
    file_sink(const std::string& n, std::ios_base::openmode m = std::ios_base::out) : name(n), mode(m) {}
};

namespace detail {
class inbuf : public std::streambuf {
public:
    inbuf() : m_out(1 << 18), m_in(1 << 18) { std::memset(&m_z, 0, sizeof(m_z)); }
    ~inbuf() { if (m_zinit) inflateEnd(&m_z); }
    void set_gzip() { m_gzip = true; }
    void set_source(std::istream* s) { m_src = s; }
protected:
    int_type underflow() override {
        if (gptr() < egptr()) return traits_type::to_int_type(*gptr());
        if (!m_src) return traits_type::eof();
        size_t got = 0;
        if (!m_gzip) {
            m_src->read(m_out.data(), m_out.size());
            got = (size_t)m_src->gcount();
        } else {
            if (!m_zinit) {
                if (inflateInit2(&m_z, 15 + 16) != Z_OK) throw std::runtime_error("inflateInit2 failed");
                m_zinit = true;
            }
            while (got == 0) {
                if (m_z.avail_in == 0) {
                    m_src->read(m_in.data(), m_in.size());
                    m_z.avail_in = (uInt)m_src->gcount();
                    m_z.next_in = (Bytef*)m_in.data();
                    if (m_z.avail_in == 0) break;
                }
                m_z.next_out = (Bytef*)m_out.data();
                m_z.avail_out = (uInt)m_out.size();
                int rc = inflate(&m_z, Z_NO_FLUSH);
                got = m_out.size() - m_z.avail_out;
                if (rc == Z_STREAM_END) { inflateReset(&m_z); }   // multi-member gzip
                else if (rc != Z_OK && rc != Z_BUF_ERROR) throw std::runtime_error("gzip error");
            }
        }
        if (got == 0) return traits_type::eof();
        setg(m_out.data(), m_out.data(), m_out.data() + got);
        return traits_type::to_int_type(*gptr());
    }
private:
    std::vector<char> m_out, m_in;
    std::istream* m_src = nullptr;
    bool m_gzip = false, m_zinit = false;
    z_stream m_z;
};

class outbuf : public std::streambuf {
public:
    outbuf() : m_buf(1 << 18), m_zout(1 << 18) {
        std::memset(&m_z, 0, sizeof(m_z));
        setp(m_buf.data(), m_buf.data() + m_buf.size());
    }
    ~outbuf() { finish(); }
    void set_gzip() { m_gzip = true; }
    void open(const std::string& name) {
        m_f = std::fopen(name.c_str(), "wb");
    }
    bool is_open() const { return m_f != nullptr; }
    void finish() {
        if (!m_f) return;
        drain(true);
        std::fclose(m_f);
        m_f = nullptr;
        if (m_zinit) { deflateEnd(&m_z); m_zinit = false; }
    }
protected:
    int_type overflow(int_type ch) override {
        drain(false);
        if (!traits_type::eq_int_type(ch, traits_type::eof())) { *pptr() = traits_type::to_char_type(ch); pbump(1); }
        return traits_type::not_eof(ch);
    }
    int sync() override { drain(false); return 0; }
private:
    void drain(bool last) {
        size_t n = pptr() - pbase();
        if (m_f) {
            if (!m_gzip) {
                if (n) std::fwrite(pbase(), 1, n, m_f);
            } else {
                if (!m_zinit) {
                    deflateInit2(&m_z, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY);
                    m_zinit = true;
                }
                m_z.next_in = (Bytef*)pbase();
                m_z.avail_in = (uInt)n;
                int rc;
                do {
                    m_z.next_out = (Bytef*)m_zout.data();
                    m_z.avail_out = (uInt)m_zout.size();
                    rc = deflate(&m_z, last ? Z_FINISH : Z_NO_FLUSH);
                    size_t have = m_zout.size() - m_z.avail_out;
                    if (have) std::fwrite(m_zout.data(), 1, have, m_f);
                } while (m_z.avail_out == 0 || (last && rc != Z_STREAM_END));
            }
        }
        setp(m_buf.data(), m_buf.data() + m_buf.size());
    }
    std::vector<char> m_buf, m_zout;
    FILE* m_f = nullptr;
    bool m_gzip = false, m_zinit = false;
    z_stream m_z;
};
}  // namespace detail

class filtering_istream : public std::istream {
public:
    filtering_istream() : std::istream(nullptr) { rdbuf(&m_buf); }
    void push(const gzip_decompressor&) { m_buf.set_gzip(); }
    void push(std::istream& src) { m_buf.set_source(&src); }
private:
    detail::inbuf m_buf;
};

class filtering_ostream : public std::ostream {
public:
    filtering_ostream() : std::ostream(nullptr) { rdbuf(&m_buf); }
    ~filtering_ostream() { m_buf.finish(); }
    void push(const gzip_compressor&, std::streamsize = 0) { m_buf.set_gzip(); }
    void push(const file_sink& s, std::streamsize = 0) { m_buf.open(s.name); }
private:
    detail::outbuf m_buf;
};

// Only referenced by the reference's dead [[deprecated]] helpers (src/file_utils.cpp:135-189, no callers).
template <class Mode>
class filtering_streambuf : public std::streambuf {
public:
    void push(const gzip_decompressor&) {}
    void push(const gzip_compressor&) {}
    template <class S> void push(S&) {}
};
template <class B, class S>
inline void copy(B&, S&) { throw std::runtime_error("boost shim: iostreams::copy is not implemented"); }
template <class B>
inline void close(B&) {}

}}  // namespace boost::iostreams
