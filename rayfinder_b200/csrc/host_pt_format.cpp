// The .pt scene container — byte-for-byte the layout written by the reference's
// serialize(OutputStream&, const PtFormat&) (pt-format/pt_format.cpp:240-269) and read by
// deserialize (pt_format.cpp:271-321):
//
//   "PTFORMAT3"                                             9 bytes, no NUL            (:238)
//   8 x  [u64 n][n * sizeof(T)]                             arrays 0..7                 (:153-160)
//   4 x  [u64 numSlices]{[u64 offsetIdx][u64 numElements]}  arrays 8..11                (:162-181)
//   1 x  [u64 n][n * u32]                                   array 12
//   [u64 numTextures] { [u32 width][u32 height][u64 numPixels][numPixels * u32 BGRA] }  (:220-225)
//
// Little-endian, packed, struct padding bytes included.  Host-only code.
#include "pt_file.h"
#include "rf_internal.h"

#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <regex>
#include <stdexcept>
#include <string>
#include <vector>

using namespace rfb200;

namespace
{
constexpr char        MAGIC[] = "PTFORMAT3";
constexpr std::size_t MAGIC_LEN = 9;
} // namespace

namespace
{
// Sequential reader over a memory range; short reads are reported instead of being ignored (the
// reference only NLRS_ASSERTs on them, pt_format.cpp:189-217, which is a no-op in release builds).
struct Reader
{
    const std::uint8_t* cur;
    const std::uint8_t* end;
    bool                ok = true;

    void read(void* dst, std::uint64_t n)
    {
        if (!ok || static_cast<std::uint64_t>(end - cur) < n)
        {
            ok = false;
            return;
        }
        std::memcpy(dst, cur, n);
        cur += n;
    }
    std::uint64_t u64()
    {
        std::uint64_t v = 0;
        read(&v, 8);
        return v;
    }
};

rf_status parse(const std::uint8_t* data, std::uint64_t size, rf_pt_file& f)
{
    Reader      r{data, data + size};
    std::string magic(MAGIC_LEN, '\0');
    r.read(magic.data(), MAGIC_LEN);
    if (!r.ok || magic != MAGIC)
    {
        // pt_format.cpp:277-291: version mismatch vs. not a .pt file at all.
        if (std::regex_search(magic, std::regex("PTFORMAT\\d")))
        {
            return setError(
                RF_ERROR_FORMAT,
                "Mismatching PtFormat file version. Invalid version in magic bytes: expected '%s', got '%s'.",
                MAGIC,
                magic.c_str());
        }
        return setError(RF_ERROR_FORMAT, "Invalid file format: expected PtFormat file.");
    }

    const auto readArray = [&](int which) {
        const std::uint64_t n = r.u64();
        if (!r.ok) return;
        const std::uint64_t bytes = n * RF_PT_ELEM_SIZE[which];
        if (n != 0 && (bytes / RF_PT_ELEM_SIZE[which] != n || static_cast<std::uint64_t>(r.end - r.cur) < bytes))
        {
            r.ok = false;
            return;
        }
        f.arrays[which].resize(bytes);
        r.read(f.arrays[which].data(), bytes);
    };
    for (int a = RF_PT_BVH_NODES; a <= RF_PT_VERTEX_INDICES; ++a) readArray(a);
    for (int a = RF_PT_MODEL_VERTEX_POSITIONS; a <= RF_PT_MODEL_VERTEX_INDICES; ++a)
    {
        readArray(a); // [u64 numSlices] then numSlices x {u64, u64} — same wire shape as an array of 16 B
        if (!r.ok) break;
        // offsetIdx + numElements <= buffer.size() (pt_format.cpp:201)
        const int           base = a - RF_PT_MODEL_VERTEX_POSITIONS + RF_PT_VERTEX_POSITIONS;
        const std::uint64_t baseCount = f.count(base);
        const auto*         s = reinterpret_cast<const std::uint64_t*>(f.arrays[a].data());
        for (std::uint64_t i = 0; i < f.count(a); ++i)
        {
            if (s[2 * i] + s[2 * i + 1] > baseCount)
            {
                return setError(RF_ERROR_FORMAT, "Invalid PtFormat file: slice %llu of array %d out of range.", (unsigned long long)i, a);
            }
        }
    }
    readArray(RF_PT_MODEL_BASE_COLOR_TEXTURE_INDICES);

    const std::uint64_t numTextures = r.u64();
    if (r.ok)
    {
        if (numTextures > static_cast<std::uint64_t>(r.end - r.cur) / 16)
        {
            r.ok = false;
        }
        else
        {
            f.textures.resize(numTextures);
            for (PtTextureData& t : f.textures)
            {
                r.read(&t.width, 4);
                r.read(&t.height, 4);
                const std::uint64_t n = r.u64();
                if (!r.ok || n > static_cast<std::uint64_t>(r.end - r.cur) / 4) // (no n * 4: it wraps for n >= 2^62)
                {
                    r.ok = false;
                    break;
                }
                // Every consumer reads width * height texels (Texture::pixels() is exactly that in the reference,
                // common/texture.hpp; reference_path_tracer.cpp:229-269 uploads pixels().size()): a file that says otherwise
                // is malformed, not a short read.
                if (n != static_cast<std::uint64_t>(t.width) * t.height)
                {
                    return setError(RF_ERROR_FORMAT, "Invalid PtFormat file: texture %llu has %llu pixels for %ux%u.",
                                    (unsigned long long)(&t - f.textures.data()), (unsigned long long)n, t.width, t.height);
                }
                t.pixels.resize(n);
                r.read(t.pixels.data(), n * 4);
            }
        }
    }
    if (!r.ok)
    {
        return setError(RF_ERROR_IO, "Unexpected end of PtFormat stream.");
    }
    return RF_OK;
}

std::uint64_t serializedSize(const rf_pt_file& f)
{
    std::uint64_t n = MAGIC_LEN;
    for (int a = 0; a < RF_PT_NUM_ARRAYS; ++a) n += 8 + f.arrays[a].size();
    n += 8;
    for (const PtTextureData& t : f.textures) n += 16 + 4 * t.pixels.size();
    return n;
}

void serializeTo(const rf_pt_file& f, std::uint8_t* dst)
{
    const auto put = [&dst](const void* src, std::uint64_t n) {
        if (n) std::memcpy(dst, src, n);
        dst += n;
    };
    put(MAGIC, MAGIC_LEN);
    for (int a = 0; a < RF_PT_NUM_ARRAYS; ++a)
    {
        const std::uint64_t n = f.count(a);
        put(&n, 8);
        put(f.arrays[a].data(), f.arrays[a].size());
    }
    const std::uint64_t numTextures = f.textures.size();
    put(&numTextures, 8);
    for (const PtTextureData& t : f.textures)
    {
        put(&t.width, 4);
        put(&t.height, 4);
        const std::uint64_t n = t.pixels.size();
        put(&n, 8);
        put(t.pixels.data(), 4 * n);
    }
}

// No C++ exception may cross the extern "C" boundary (std::terminate): allocation failures of the containers
// become a status code.
template<class F>
rf_status guarded(const char* what, F&& body)
{
    try
    {
        return body();
    }
    catch (const std::bad_alloc&)
    {
        return setError(RF_ERROR_IO, "%s: out of memory", what);
    }
    catch (const std::length_error&)
    {
        return setError(RF_ERROR_INVALID_ARGUMENT, "%s: size out of range", what);
    }
    catch (const std::exception& e)
    {
        return setError(RF_ERROR_IO, "%s: %s", what, e.what());
    }
}
} // namespace

extern "C" rf_status rf_pt_create(rf_pt_file** out)
{
    if (!out) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_create: null argument");
    return guarded("rf_pt_create", [&]() -> rf_status {
        *out = new rf_pt_file();
        return RF_OK;
    });
}

extern "C" void rf_pt_destroy(rf_pt_file* f) { delete f; }

extern "C" rf_status rf_pt_load_memory(const void* data, std::uint64_t size, rf_pt_file** out)
{
    if (!data || !out) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_load_memory: null argument");
    return guarded("rf_pt_load_memory", [&]() -> rf_status {
        auto            f = std::make_unique<rf_pt_file>();
        const rf_status st = parse(static_cast<const std::uint8_t*>(data), size, *f);
        if (st != RF_OK) return st;
        *out = f.release();
        return RF_OK;
    });
}

extern "C" rf_status rf_pt_load(const char* path, rf_pt_file** out)
{
    if (!path || !out) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_load: null argument");
    std::FILE* fp = std::fopen(path, "rb");
    if (!fp)
    {
        return setError(RF_ERROR_IO, "Failed to open file: %s", path); // common/file_stream.cpp:13-16
    }
    return guarded("rf_pt_load", [&]() -> rf_status {
        std::fseek(fp, 0, SEEK_END);
        const long size = std::ftell(fp);
        std::fseek(fp, 0, SEEK_SET);
        std::vector<std::uint8_t> buf;
        try
        {
            buf.resize(size > 0 ? static_cast<std::size_t>(size) : 0);
        }
        catch (...)
        {
            std::fclose(fp);
            throw;
        }
        const std::size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), fp);
        std::fclose(fp);
        if (got != buf.size()) return setError(RF_ERROR_IO, "Failed to read file: %s", path);
        return rf_pt_load_memory(buf.data(), buf.size(), out);
    });
}

extern "C" rf_status rf_pt_save_memory(const rf_pt_file* f, void* dst, std::uint64_t capacity, std::uint64_t* size)
{
    if (!f || !size) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_save_memory: null argument");
    *size = serializedSize(*f);
    if (!dst) return RF_OK; // size query
    if (capacity < *size) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_save_memory: buffer too small");
    serializeTo(*f, static_cast<std::uint8_t*>(dst));
    return RF_OK;
}

extern "C" rf_status rf_pt_save(const rf_pt_file* f, const char* path)
{
    if (!f || !path) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_save: null argument");
    return guarded("rf_pt_save", [&]() -> rf_status {
        std::vector<std::uint8_t> buf(serializedSize(*f));
        serializeTo(*f, buf.data());
        std::FILE* fp = std::fopen(path, "wb");
        if (!fp) return setError(RF_ERROR_IO, "Failed to open file: %s", path);
        const std::size_t put = std::fwrite(buf.data(), 1, buf.size(), fp);
        std::fclose(fp);
        if (put != buf.size()) return setError(RF_ERROR_IO, "Failed to write file: %s", path);
        return RF_OK;
    });
}

extern "C" rf_status rf_pt_array(const rf_pt_file* f, int32_t which, const void** data, std::uint64_t* count, std::uint64_t* elem_size)
{
    if (!f || which < 0 || which >= RF_PT_NUM_ARRAYS) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_array: bad argument");
    if (data) *data = f->arrays[which].data();
    if (count) *count = f->count(which);
    if (elem_size) *elem_size = RF_PT_ELEM_SIZE[which];
    return RF_OK;
}

extern "C" rf_status rf_pt_set_array(rf_pt_file* f, int32_t which, const void* data, std::uint64_t count)
{
    if (!f || which < 0 || which >= RF_PT_NUM_ARRAYS || (!data && count)) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_set_array: bad argument");
    if (count > (~0ull >> 1) / RF_PT_ELEM_SIZE[which]) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_set_array: count out of range");
    return guarded("rf_pt_set_array", [&]() -> rf_status {
        const auto* p = static_cast<const std::uint8_t*>(data);
        f->arrays[which].assign(p, p + count * RF_PT_ELEM_SIZE[which]);
        return RF_OK;
    });
}

extern "C" std::uint64_t rf_pt_num_textures(const rf_pt_file* f) { return f ? f->textures.size() : 0; }

extern "C" rf_status rf_pt_texture(const rf_pt_file* f, std::uint64_t idx, rf_texture* out)
{
    if (!f || !out || idx >= f->textures.size()) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_texture: bad argument");
    const PtTextureData& t = f->textures[idx];
    *out = rf_texture{t.pixels.data(), t.width, t.height};
    return RF_OK;
}

extern "C" rf_status rf_pt_add_texture(rf_pt_file* f, const std::uint32_t* pixels, std::uint32_t width, std::uint32_t height)
{
    if (!f || !pixels) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_add_texture: null argument");
    return guarded("rf_pt_add_texture", [&]() -> rf_status {
        PtTextureData t;
        t.width = width, t.height = height;
        t.pixels.assign(pixels, pixels + static_cast<std::uint64_t>(width) * height);
        f->textures.push_back(std::move(t));
        return RF_OK;
    });
}

extern "C" rf_status rf_pt_scene(const rf_pt_file* f, rf_scene* out, rf_texture* textures)
{
    if (!f || !out || (!textures && !f->textures.empty())) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pt_scene: null argument");
    // The four spans handed to the renderer, pt/main.cpp:150-155.
    out->bvh_nodes = reinterpret_cast<const rf_bvh_node*>(f->arrays[RF_PT_BVH_NODES].data());
    out->num_bvh_nodes = f->count(RF_PT_BVH_NODES);
    out->position_attributes = reinterpret_cast<const rf_position_attribute*>(f->arrays[RF_PT_TRIANGLE_POSITION_ATTRIBUTES].data());
    out->num_position_attributes = f->count(RF_PT_TRIANGLE_POSITION_ATTRIBUTES);
    out->vertex_attributes = reinterpret_cast<const rf_vertex_attributes*>(f->arrays[RF_PT_TRIANGLE_VERTEX_ATTRIBUTES].data());
    out->num_vertex_attributes = f->count(RF_PT_TRIANGLE_VERTEX_ATTRIBUTES);
    for (std::size_t i = 0; i < f->textures.size(); ++i)
    {
        textures[i] = rf_texture{f->textures[i].pixels.data(), f->textures[i].width, f->textures[i].height};
    }
    out->base_color_textures = textures;
    out->num_base_color_textures = f->textures.size();
    return RF_OK;
}
