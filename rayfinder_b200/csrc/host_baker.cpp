// Scene baker: nlrs::PtFormat(std::filesystem::path gltfPath) (pt-format/pt_format.cpp:20-151) on the host, i.e.
//   GltfModel(gltfPath)      common/gltf_model.cpp:266-465   parse, node transforms, attributes, base-colour textures, mesh sort
//   FlattenedModel(model)    common/flattened_model.cpp:8-46  one Positions / Normals / TexCoords / texture index per triangle
//   buildBvh + reorder       common/bvh.cpp:263-291, bvh.hpp:36-46 (rf_build_bvh, byte-identical to the reference's)
//   the rasteriser's arrays  pt_format.cpp:84-148
// — the step before the render path.  The reference's parsers and maths are un-vendored third-party
// code (cgltf 1.13, glm 0.9.9.8, stb_image); what they compute is restated here:
//   * glTF 2.0 JSON / GLB container: own parser, the subset cgltf hands to gltf_model.cpp (scenes, nodes, meshes, accessors,
//     buffer views, buffers incl. data: URIs and external files, materials, textures, samplers, images);
//   * glm: column-major fp32 formulas in glm's operation order — mat4*mat4, mat4*vec4, scale/translate/mat4_cast,
//     inverseTranspose by cofactors, normalize(vec4) with the w component included (gltf_model.cpp:427-428: the reference
//     normalises the 4-vector, so a node with a translation yields non-unit xyz normals — reproduced);
//   * stb_image: host_image.cpp.
// Meshes are ordered with std::sort and the reference's comparator (gltf_model.cpp:462): with libstdc++ that is the order a
// Linux build of the reference produces (equal keys are NOT kept in file order beyond 16 meshes).
#include "pt_file.h"
#include "rf_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fs = std::filesystem;

namespace rfb200
{
bool decodeImageRgba8(const std::uint8_t* data, std::size_t size, std::vector<std::uint8_t>& rgba, std::uint32_t& width, std::uint32_t& height, std::string& error);

namespace
{
// ---------------------------------------------------------------------------------------------- JSON
struct Json
{
    enum Type
    {
        Null,
        Bool,
        Number,
        String,
        Array,
        Object
    } type = Null;
    bool                                      boolean = false;
    std::string                               text; // String: the value; Number: its source text
    std::vector<Json>                         items;
    std::vector<std::pair<std::string, Json>> members;

    const Json* find(const char* key) const
    {
        for (const auto& m : members)
            if (m.first == key) return &m.second;
        return nullptr;
    }
    bool        has(const char* key) const { return find(key) != nullptr; }
    // cgltf_json_to_float: (float)atof(token); cgltf_json_to_int / _size: atoi / atoll
    float       asFloat() const { return static_cast<float>(std::atof(text.c_str())); }
    long long   asInt() const { return std::atoll(text.c_str()); }
    float       floatAt(const char* key, float fallback) const { const Json* v = find(key); return v ? v->asFloat() : fallback; }
    long long   intAt(const char* key, long long fallback) const { const Json* v = find(key); return v ? v->asInt() : fallback; }
    std::size_t size() const { return items.size(); }
};

struct ParseError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

class JsonParser
{
public:
    JsonParser(const char* begin, const char* end) : p(begin), end(end) {}
    Json parse()
    {
        Json v = value();
        skipSpace();
        return v;
    }

private:
    const char* p;
    const char* end;
    void        skipSpace()
    {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    }
    [[noreturn]] void bad() { throw ParseError("invalid json"); }
    Json              value()
    {
        skipSpace();
        if (p >= end) bad();
        Json v;
        const char c = *p;
        if (c == '{')
        {
            v.type = Json::Object;
            ++p;
            skipSpace();
            if (p < end && *p == '}')
            {
                ++p;
                return v;
            }
            for (;;)
            {
                skipSpace();
                if (p >= end || *p != '"') bad();
                std::string key = string();
                skipSpace();
                if (p >= end || *p != ':') bad();
                ++p;
                v.members.emplace_back(std::move(key), value());
                skipSpace();
                if (p < end && *p == ',')
                {
                    ++p;
                    continue;
                }
                if (p < end && *p == '}')
                {
                    ++p;
                    return v;
                }
                bad();
            }
        }
        if (c == '[')
        {
            v.type = Json::Array;
            ++p;
            skipSpace();
            if (p < end && *p == ']')
            {
                ++p;
                return v;
            }
            for (;;)
            {
                v.items.push_back(value());
                skipSpace();
                if (p < end && *p == ',')
                {
                    ++p;
                    continue;
                }
                if (p < end && *p == ']')
                {
                    ++p;
                    return v;
                }
                bad();
            }
        }
        if (c == '"')
        {
            v.type = Json::String;
            v.text = string();
            return v;
        }
        if (end - p >= 4 && std::strncmp(p, "true", 4) == 0)
        {
            v.type = Json::Bool, v.boolean = true, p += 4;
            return v;
        }
        if (end - p >= 5 && std::strncmp(p, "false", 5) == 0)
        {
            v.type = Json::Bool, p += 5;
            return v;
        }
        if (end - p >= 4 && std::strncmp(p, "null", 4) == 0)
        {
            p += 4;
            return v;
        }
        if (c == '-' || (c >= '0' && c <= '9'))
        {
            const char* start = p;
            while (p < end && (*p == '-' || *p == '+' || *p == '.' || *p == 'e' || *p == 'E' || (*p >= '0' && *p <= '9'))) ++p;
            v.type = Json::Number;
            v.text.assign(start, p);
            return v;
        }
        bad();
    }
    std::string string()
    {
        std::string out;
        ++p; // opening quote
        while (p < end && *p != '"')
        {
            if (*p == '\\')
            {
                if (++p >= end) bad();
                switch (*p)
                {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u':
                {
                    if (end - p < 5) bad();
                    const unsigned cp = static_cast<unsigned>(std::strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16));
                    p += 4;
                    if (cp < 0x80) out += static_cast<char>(cp);
                    else if (cp < 0x800) out += static_cast<char>(0xC0 | (cp >> 6)), out += static_cast<char>(0x80 | (cp & 0x3F));
                    else out += static_cast<char>(0xE0 | (cp >> 12)), out += static_cast<char>(0x80 | ((cp >> 6) & 0x3F)), out += static_cast<char>(0x80 | (cp & 0x3F));
                    break;
                }
                default: out += *p; break; // \" \\ \/
                }
                ++p;
            }
            else
            {
                out += *p++;
            }
        }
        if (p >= end) bad();
        ++p;
        return out;
    }
};

// ---------------------------------------------------------------------------------------------- glm
struct Vec4
{
    float x, y, z, w;
};
inline Vec4 operator*(Vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline Vec4 operator+(Vec4 a, Vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
struct Mat4
{
    Vec4 c[4]; // columns
    static Mat4 identity() { return {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}}; }
    float       at(int col, int row) const { return (&c[col].x)[row]; }
    float&      at(int col, int row) { return (&c[col].x)[row]; }
};
// operator*(mat4, mat4), glm/detail/type_mat4x4.inl: Result[i] = A0*B[i][0] + A1*B[i][1] + A2*B[i][2] + A3*B[i][3]
Mat4 mul(const Mat4& a, const Mat4& b)
{
    Mat4 r;
    for (int i = 0; i < 4; ++i) r.c[i] = ((a.c[0] * b.c[i].x + a.c[1] * b.c[i].y) + a.c[2] * b.c[i].z) + a.c[3] * b.c[i].w;
    return r;
}
// operator*(mat4, vec4): (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
Vec4 mul(const Mat4& m, Vec4 v) { return (m.c[0] * v.x + m.c[1] * v.y) + (m.c[2] * v.z + m.c[3] * v.w); }

// glm::inverseTranspose(mat4), glm/gtc/matrix_inverse.inl
Mat4 inverseTranspose(const Mat4& m)
{
    const auto M = [&m](int col, int row) { return m.at(col, row); };
    const float SubFactor00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    const float SubFactor01 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const float SubFactor02 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    const float SubFactor03 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    const float SubFactor04 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    const float SubFactor05 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    const float SubFactor06 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    const float SubFactor07 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    const float SubFactor08 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    const float SubFactor09 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    const float SubFactor10 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    const float SubFactor11 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    const float SubFactor12 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    const float SubFactor13 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    const float SubFactor14 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    const float SubFactor15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    const float SubFactor16 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    const float SubFactor17 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);

    Mat4 inv;
    inv.at(0, 0) = +(M(1, 1) * SubFactor00 - M(1, 2) * SubFactor01 + M(1, 3) * SubFactor02);
    inv.at(0, 1) = -(M(1, 0) * SubFactor00 - M(1, 2) * SubFactor03 + M(1, 3) * SubFactor04);
    inv.at(0, 2) = +(M(1, 0) * SubFactor01 - M(1, 1) * SubFactor03 + M(1, 3) * SubFactor05);
    inv.at(0, 3) = -(M(1, 0) * SubFactor02 - M(1, 1) * SubFactor04 + M(1, 2) * SubFactor05);

    inv.at(1, 0) = -(M(0, 1) * SubFactor00 - M(0, 2) * SubFactor01 + M(0, 3) * SubFactor02);
    inv.at(1, 1) = +(M(0, 0) * SubFactor00 - M(0, 2) * SubFactor03 + M(0, 3) * SubFactor04);
    inv.at(1, 2) = -(M(0, 0) * SubFactor01 - M(0, 1) * SubFactor03 + M(0, 3) * SubFactor05);
    inv.at(1, 3) = +(M(0, 0) * SubFactor02 - M(0, 1) * SubFactor04 + M(0, 2) * SubFactor05);

    inv.at(2, 0) = +(M(0, 1) * SubFactor06 - M(0, 2) * SubFactor07 + M(0, 3) * SubFactor08);
    inv.at(2, 1) = -(M(0, 0) * SubFactor06 - M(0, 2) * SubFactor09 + M(0, 3) * SubFactor10);
    inv.at(2, 2) = +(M(0, 0) * SubFactor07 - M(0, 1) * SubFactor09 + M(0, 3) * SubFactor11);
    inv.at(2, 3) = -(M(0, 0) * SubFactor08 - M(0, 1) * SubFactor10 + M(0, 2) * SubFactor11);

    inv.at(3, 0) = -(M(0, 1) * SubFactor12 - M(0, 2) * SubFactor13 + M(0, 3) * SubFactor14);
    inv.at(3, 1) = +(M(0, 0) * SubFactor12 - M(0, 2) * SubFactor15 + M(0, 3) * SubFactor16);
    inv.at(3, 2) = -(M(0, 0) * SubFactor13 - M(0, 1) * SubFactor15 + M(0, 3) * SubFactor17);
    inv.at(3, 3) = +(M(0, 0) * SubFactor14 - M(0, 1) * SubFactor16 + M(0, 2) * SubFactor17);

    const float determinant = +M(0, 0) * inv.at(0, 0) + M(0, 1) * inv.at(0, 1) + M(0, 2) * inv.at(0, 2) + M(0, 3) * inv.at(0, 3);
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) inv.at(col, row) = inv.at(col, row) / determinant;
    return inv;
}

// glm::normalize(vec4) = v * inversesqrt(dot(v, v)), dot = (x*x + y*y) + (z*z + w*w), inversesqrt(x) = 1 / sqrt(x)
Vec4 normalize4(Vec4 v)
{
    const float d = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    return v * (1.0f / std::sqrt(d));
}

// ---------------------------------------------------------------------------------------------- glTF
std::vector<std::uint8_t> readFile(const fs::path& path)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path.string());
    const std::streamsize     n = f.tellg();
    std::vector<std::uint8_t> data(n > 0 ? static_cast<std::size_t>(n) : 0);
    f.seekg(0);
    if (!data.empty()) f.read(reinterpret_cast<char*>(data.data()), n);
    if (!f) throw std::runtime_error("cannot read " + path.string());
    return data;
}

std::vector<std::uint8_t> decodeBase64(const std::string& text, std::size_t begin)
{
    std::vector<std::uint8_t> out;
    unsigned                  buffer = 0;
    int                       bits = 0;
    for (std::size_t i = begin; i < text.size(); ++i)
    {
        const char c = text[i];
        int        v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+') v = 62;
        else if (c == '/') v = 63;
        else break; // '=' padding
        buffer = (buffer << 6) | static_cast<unsigned>(v);
        bits += 6;
        if (bits >= 8)
        {
            bits -= 8;
            out.push_back(static_cast<std::uint8_t>((buffer >> bits) & 0xFF));
        }
    }
    return out;
}

struct Gltf
{
    Json                                   json;
    std::vector<std::vector<std::uint8_t>> buffers; // one per json.buffers entry
    fs::path                               path;

    const Json& array(const char* key) const
    {
        static const Json empty;
        const Json*       v = json.find(key);
        return v ? *v : empty;
    }
    // cgltf_buffer_view_data
    std::pair<const std::uint8_t*, std::size_t> view(std::size_t index) const
    {
        const Json&       bv = array("bufferViews").items.at(index);
        const std::size_t buffer = static_cast<std::size_t>(bv.intAt("buffer", 0));
        const std::size_t offset = static_cast<std::size_t>(bv.intAt("byteOffset", 0));
        const std::size_t length = static_cast<std::size_t>(bv.intAt("byteLength", 0));
        const auto&       data = buffers.at(buffer);
        if (offset > data.size() || length > data.size() - offset) throw std::runtime_error("buffer view out of range");
        return {data.data() + offset, length};
    }
};

struct Accessor
{
    const std::uint8_t* base = nullptr;
    std::size_t         count = 0, stride = 0, elementSize = 0;
    int                 componentType = 0, numComponents = 0;
};

Accessor accessorOf(const Gltf& g, std::size_t index)
{
    const Json& a = g.array("accessors").items.at(index);
    if (a.has("sparse")) throw std::runtime_error("sparse accessors are not supported");
    Accessor acc;
    acc.componentType = static_cast<int>(a.intAt("componentType", 0));
    acc.count = static_cast<std::size_t>(a.intAt("count", 0));
    const Json* type = a.find("type");
    const std::string t = type ? type->text : "";
    acc.numComponents = t == "SCALAR" ? 1 : t == "VEC2" ? 2 : t == "VEC3" ? 3 : t == "VEC4" ? 4 : t == "MAT4" ? 16 : 0;
    const int componentSize = (acc.componentType == 5120 || acc.componentType == 5121) ? 1 : (acc.componentType == 5122 || acc.componentType == 5123) ? 2 : 4;
    acc.elementSize = static_cast<std::size_t>(componentSize) * acc.numComponents;
    const std::size_t viewIndex = static_cast<std::size_t>(a.intAt("bufferView", -1));
    const Json&       bv = g.array("bufferViews").items.at(viewIndex);
    const auto [data, length] = g.view(viewIndex);
    const std::size_t offset = static_cast<std::size_t>(a.intAt("byteOffset", 0));
    const std::size_t byteStride = static_cast<std::size_t>(bv.intAt("byteStride", 0));
    acc.stride = byteStride ? byteStride : acc.elementSize; // cgltf: accessor->stride
    if (acc.count && (offset > length || (acc.count - 1) * acc.stride + acc.elementSize > length - offset)) throw std::runtime_error("accessor out of range");
    acc.base = data + offset;
    return acc;
}

// cgltf_accessor_read_uint for the component types an index accessor may have
std::uint32_t readIndex(const Accessor& a, std::size_t i)
{
    const std::uint8_t* p = a.base + i * a.stride;
    switch (a.componentType)
    {
    case 5121: return *p;
    case 5123:
    {
        std::uint16_t v;
        std::memcpy(&v, p, 2);
        return v;
    }
    case 5125:
    {
        std::uint32_t v;
        std::memcpy(&v, p, 4);
        return v;
    }
    default: throw std::runtime_error("unsupported index component type");
    }
}

struct Mesh // nlrs::GltfMesh
{
    std::vector<float>         positions, normals; // 3 per vertex
    std::vector<float>         texCoords;          // 2 per vertex
    std::vector<std::uint32_t> indices;
    std::size_t                baseColorTextureIndex = 0;
};

struct Model // nlrs::GltfModel
{
    std::vector<Mesh>          meshes;
    std::vector<PtTextureData> textures;
};

std::uint32_t fnv1a(const void* data, std::size_t size) // gltf_model.cpp:123-135
{
    const auto*   p = static_cast<const unsigned char*>(data);
    std::uint32_t hash = 2166136261u;
    for (std::size_t i = 0; i < size; ++i)
    {
        hash ^= p[i];
        hash *= 16777619u;
    }
    return hash;
}

PtTextureData textureFromMemory(const std::uint8_t* data, std::size_t size)
{
    std::vector<std::uint8_t> rgba;
    std::string               error;
    PtTextureData             t;
    if (!decodeImageRgba8(data, size, rgba, t.width, t.height, error)) throw std::runtime_error("Failed to decode image: " + error);
    t.pixels.resize(static_cast<std::size_t>(t.width) * t.height);
    for (std::size_t i = 0; i < t.pixels.size(); ++i)
    {
        const std::uint32_t r = rgba[4 * i], g = rgba[4 * i + 1], b = rgba[4 * i + 2];
        t.pixels[i] = b | (g << 8) | (r << 16) | (255u << 24); // texture.cpp:41-47
    }
    return t;
}

Gltf loadGltf(const fs::path& gltfPath)
{
    if (!fs::exists(gltfPath)) throw std::runtime_error("The gltf file " + gltfPath.string() + " does not exist.");
    Gltf g;
    g.path = gltfPath;
    std::vector<std::uint8_t> glbBinary;
    bool                      haveGlbBinary = false;
    try
    {
        const std::vector<std::uint8_t> file = readFile(gltfPath);
        const char*                     jsonBegin = reinterpret_cast<const char*>(file.data());
        const char*                     jsonEnd = jsonBegin + file.size();
        if (file.size() >= 12 && std::memcmp(file.data(), "glTF", 4) == 0)
        {
            std::uint32_t version, length;
            std::memcpy(&version, file.data() + 4, 4);
            std::memcpy(&length, file.data() + 8, 4);
            if (version != 2 || length > file.size()) throw ParseError("bad glb header");
            std::size_t pos = 12;
            jsonBegin = jsonEnd = nullptr;
            while (pos + 8 <= length)
            {
                std::uint32_t chunkLength, chunkType;
                std::memcpy(&chunkLength, file.data() + pos, 4);
                std::memcpy(&chunkType, file.data() + pos + 4, 4);
                if (pos + 8 + chunkLength > length) throw ParseError("bad glb chunk");
                if (chunkType == 0x4E4F534Au && !jsonBegin)
                {
                    jsonBegin = reinterpret_cast<const char*>(file.data() + pos + 8);
                    jsonEnd = jsonBegin + chunkLength;
                }
                else if (chunkType == 0x004E4942u && !haveGlbBinary)
                {
                    glbBinary.assign(file.data() + pos + 8, file.data() + pos + 8 + chunkLength);
                    haveGlbBinary = true;
                }
                pos += 8 + chunkLength;
            }
            if (!jsonBegin) throw ParseError("no json chunk");
        }
        g.json = JsonParser(jsonBegin, jsonEnd).parse();
        if (g.json.type != Json::Object) throw ParseError("not an object");
    }
    catch (const std::exception&)
    {
        throw std::runtime_error("Failed to parse gltf file " + gltfPath.string() + ".");
    }
    // cgltf_load_buffers
    try
    {
        const Json& buffers = g.array("buffers");
        for (std::size_t i = 0; i < buffers.size(); ++i)
        {
            const Json& uri = buffers.items[i].has("uri") ? *buffers.items[i].find("uri") : Json{};
            if (uri.type != Json::String)
            {
                if (i != 0 || !haveGlbBinary) throw std::runtime_error("buffer without data");
                g.buffers.push_back(std::move(glbBinary));
            }
            else if (uri.text.rfind("data:", 0) == 0)
            {
                const std::size_t comma = uri.text.find(',');
                if (comma == std::string::npos || uri.text.find(";base64") == std::string::npos) throw std::runtime_error("unknown data uri");
                g.buffers.push_back(decodeBase64(uri.text, comma + 1));
            }
            else
            {
                g.buffers.push_back(readFile(gltfPath.parent_path() / uri.text));
            }
            const std::size_t byteLength = static_cast<std::size_t>(buffers.items[i].intAt("byteLength", 0));
            if (g.buffers.back().size() < byteLength) throw std::runtime_error("buffer too short");
        }
    }
    catch (const std::exception&)
    {
        throw std::runtime_error("Failed to load gltf buffers for " + gltfPath.string() + ".");
    }
    return g;
}

// traverseNodeHierarchy, gltf_model.cpp:28-73
void traverse(const Gltf& g, std::size_t nodeIndex, const Mat4& parent, std::vector<std::pair<Mat4, Mat4>>& transforms)
{
    const Json& node = g.array("nodes").items.at(nodeIndex);
    Mat4        local;
    if (const Json* m = node.find("matrix"))
    {
        if (m->size() != 16) throw std::runtime_error("node matrix needs 16 values");
        for (int i = 0; i < 16; ++i) (&local.c[0].x)[i] = m->items[i].asFloat();
    }
    else
    {
        float       t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
        const auto  fill = [&node](const char* key, float* dst, std::size_t n) {
            if (const Json* v = node.find(key))
                for (std::size_t i = 0; i < n && i < v->size(); ++i) dst[i] = v->items[i].asFloat();
        };
        fill("translation", t, 3), fill("rotation", q, 4), fill("scale", s, 3);
        const Mat4 one = Mat4::identity();
        // glm::scale(mat4(1), s)
        Mat4 scale;
        scale.c[0] = one.c[0] * s[0], scale.c[1] = one.c[1] * s[1], scale.c[2] = one.c[2] * s[2], scale.c[3] = one.c[3];
        // glm::toMat4(q) = mat4(mat3_cast(q)), q = (x, y, z, w)
        const float qx = q[0], qy = q[1], qz = q[2], qw = q[3];
        const float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
        Mat4        rotation = one;
        rotation.at(0, 0) = 1.0f - 2.0f * (qyy + qzz);
        rotation.at(0, 1) = 2.0f * (qxy + qwz);
        rotation.at(0, 2) = 2.0f * (qxz - qwy);
        rotation.at(1, 0) = 2.0f * (qxy - qwz);
        rotation.at(1, 1) = 1.0f - 2.0f * (qxx + qzz);
        rotation.at(1, 2) = 2.0f * (qyz + qwx);
        rotation.at(2, 0) = 2.0f * (qxz + qwy);
        rotation.at(2, 1) = 2.0f * (qyz - qwx);
        rotation.at(2, 2) = 1.0f - 2.0f * (qxx + qyy);
        // glm::translate(mat4(1), t): Result[3] = m[0]*t[0] + m[1]*t[1] + m[2]*t[2] + m[3]
        Mat4 translation = one;
        translation.c[3] = ((one.c[0] * t[0] + one.c[1] * t[1]) + one.c[2] * t[2]) + one.c[3];
        local = mul(mul(translation, rotation), scale);
    }
    const Mat4 transform = mul(parent, local);
    const Mat4 normalMatrix = inverseTranspose(transform);
    if (const Json* mesh = node.find("mesh"))
    {
        const std::size_t meshIndex = static_cast<std::size_t>(mesh->asInt());
        if (meshIndex < transforms.size()) transforms[meshIndex] = {transform, normalMatrix};
    }
    if (const Json* children = node.find("children"))
        for (const Json& child : children->items) traverse(g, static_cast<std::size_t>(child.asInt()), transform, transforms);
}

Model loadModel(const fs::path& gltfPath)
{
    const Gltf  g = loadGltf(gltfPath);
    const Json& meshes = g.array("meshes");
    std::vector<std::pair<Mat4, Mat4>> transforms(meshes.size(), {Mat4::identity(), Mat4::identity()});
    {
        const Json& scenes = g.array("scenes");
        if (scenes.size() != 1) throw std::runtime_error("expected exactly one scene"); // NLRS_ASSERT(data->scenes_count == 1)
        const std::size_t sceneIndex = static_cast<std::size_t>(g.json.intAt("scene", 0));
        const Json*       roots = scenes.items.at(sceneIndex < scenes.size() ? sceneIndex : 0).find("nodes");
        if (roots)
            for (const Json& root : roots->items) traverse(g, static_cast<std::size_t>(root.asInt()), Mat4::identity(), transforms);
    }

    Model model;
    // BaseColorTextureBuilder, gltf_model.cpp:143-264
    struct ImageLookup
    {
        std::size_t image, texture;
    };
    struct FactorLookup
    {
        std::uint32_t hash;
        std::size_t   texture;
    };
    std::vector<ImageLookup>  imageLookups;
    std::vector<FactorLookup> factorLookups;
    const auto                addBaseColor = [&](const Json& pbr) -> std::size_t {
        const Json* baseColorTexture = pbr.find("baseColorTexture");
        if (baseColorTexture && baseColorTexture->has("index"))
        {
            const Json& texture = g.array("textures").items.at(static_cast<std::size_t>(baseColorTexture->intAt("index", 0)));
            if (const Json* sampler = texture.find("sampler"))
            {
                const Json& s = g.array("samplers").items.at(static_cast<std::size_t>(sampler->asInt()));
                if (s.intAt("wrapS", 10497) != 10497 || s.intAt("wrapT", 10497) != 10497) throw std::runtime_error("only REPEAT samplers are supported");
            }
            if (!texture.has("source")) throw std::runtime_error("texture without an image");
            const std::size_t imageIndex = static_cast<std::size_t>(texture.intAt("source", 0));
            for (const ImageLookup& l : imageLookups)
                if (l.image == imageIndex) return l.texture;
            const std::size_t textureIndex = model.textures.size();
            imageLookups.push_back({imageIndex, textureIndex});
            const Json& image = g.array("images").items.at(imageIndex);
            if (const Json* view = image.find("bufferView"))
            {
                const auto [data, length] = g.view(static_cast<std::size_t>(view->asInt()));
                model.textures.push_back(textureFromMemory(data, length));
            }
            else
            {
                const Json* uri = image.find("uri");
                if (!uri) throw std::runtime_error("image without data");
                if (uri->text.rfind("data:", 0) == 0)
                {
                    const std::vector<std::uint8_t> bytes = decodeBase64(uri->text, uri->text.find(',') + 1);
                    model.textures.push_back(textureFromMemory(bytes.data(), bytes.size()));
                }
                else
                {
                    const fs::path imagePath = gltfPath.parent_path() / uri->text;
                    if (!fs::exists(imagePath)) throw std::runtime_error("The image " + imagePath.string() + " does not exist.");
                    const std::vector<std::uint8_t> bytes = readFile(imagePath);
                    model.textures.push_back(textureFromMemory(bytes.data(), bytes.size()));
                }
            }
            return textureIndex;
        }
        float factor[4] = {1.0f, 1.0f, 1.0f, 1.0f}; // cgltf's default base_color_factor
        if (const Json* f = pbr.find("baseColorFactor"))
            for (std::size_t i = 0; i < 4 && i < f->size(); ++i) factor[i] = f->items[i].asFloat();
        const std::uint32_t hash = fnv1a(factor, sizeof(factor));
        for (const FactorLookup& l : factorLookups)
            if (l.hash == hash) return l.texture;
        const std::size_t textureIndex = model.textures.size();
        factorLookups.push_back({hash, textureIndex});
        // Texture::fromPixel, texture.cpp:54-65
        const std::uint32_t r8 = static_cast<std::uint32_t>(factor[0] * 255.0f), g8 = static_cast<std::uint32_t>(factor[1] * 255.0f);
        const std::uint32_t b8 = static_cast<std::uint32_t>(factor[2] * 255.0f), a8 = static_cast<std::uint32_t>(factor[3] * 255.0f);
        PtTextureData       t;
        t.width = t.height = 1;
        t.pixels = {b8 | (g8 << 8) | (r8 << 16) | (a8 << 24)};
        model.textures.push_back(std::move(t));
        return textureIndex;
    };

    for (std::size_t meshIndex = 0; meshIndex < meshes.size(); ++meshIndex)
    {
        const Json* primitives = meshes.items[meshIndex].find("primitives");
        if (!primitives) continue;
        for (const Json& primitive : primitives->items)
        {
            if (primitive.intAt("mode", 4) != 4) throw std::runtime_error("only triangle primitives are supported");
            Mesh mesh;
            // material
            {
                if (!primitive.has("material")) throw std::runtime_error("primitive without a material");
                const Json& material = g.array("materials").items.at(static_cast<std::size_t>(primitive.intAt("material", 0)));
                const Json* pbr = material.find("pbrMetallicRoughness");
                if (!pbr) throw std::runtime_error("material without pbrMetallicRoughness");
                mesh.baseColorTextureIndex = addBaseColor(*pbr);
            }
            // indices
            {
                if (!primitive.has("indices")) throw std::runtime_error("primitive without indices");
                const Accessor a = accessorOf(g, static_cast<std::size_t>(primitive.intAt("indices", 0)));
                if (a.numComponents != 1 || a.count % 3 != 0) throw std::runtime_error("index accessor must hold whole triangles");
                mesh.indices.resize(a.count);
                for (std::size_t i = 0; i < a.count; ++i) mesh.indices[i] = readIndex(a, i);
            }
            // attributes
            {
                const Json* attributes = primitive.find("attributes");
                if (!attributes || !attributes->has("POSITION") || !attributes->has("NORMAL") || !attributes->has("TEXCOORD_0"))
                    throw std::runtime_error("primitive needs POSITION, NORMAL and TEXCOORD_0");
                const Accessor p = accessorOf(g, static_cast<std::size_t>(attributes->intAt("POSITION", 0)));
                const Accessor n = accessorOf(g, static_cast<std::size_t>(attributes->intAt("NORMAL", 0)));
                const Accessor t = accessorOf(g, static_cast<std::size_t>(attributes->intAt("TEXCOORD_0", 0)));
                if (p.componentType != 5126 || p.numComponents != 3 || n.componentType != 5126 || n.numComponents != 3 || t.componentType != 5126 ||
                    t.numComponents != 2 || p.count != n.count || p.count != t.count)
                    throw std::runtime_error("attributes must be float vec3 / vec3 / vec2 of equal length");
                const auto& [transform, normalMatrix] = transforms[meshIndex];
                mesh.positions.resize(3 * p.count), mesh.normals.resize(3 * p.count), mesh.texCoords.resize(2 * p.count);
                for (std::size_t i = 0; i < p.count; ++i)
                {
                    float local[3];
                    std::memcpy(local, p.base + i * p.stride, 12);
                    const Vec4 world = mul(transform, Vec4{local[0], local[1], local[2], 1.0f});
                    mesh.positions[3 * i] = world.x, mesh.positions[3 * i + 1] = world.y, mesh.positions[3 * i + 2] = world.z;
                    std::memcpy(local, n.base + i * n.stride, 12);
                    const Vec4 normal = normalize4(mul(normalMatrix, Vec4{local[0], local[1], local[2], 0.0f}));
                    mesh.normals[3 * i] = normal.x, mesh.normals[3 * i + 1] = normal.y, mesh.normals[3 * i + 2] = normal.z;
                    std::memcpy(&mesh.texCoords[2 * i], t.base + i * t.stride, 8);
                }
                for (const std::uint32_t index : mesh.indices)
                    if (index >= p.count) throw std::runtime_error("vertex index out of range");
            }
            model.meshes.push_back(std::move(mesh));
        }
    }
    std::sort(model.meshes.begin(), model.meshes.end(), [](const Mesh& a, const Mesh& b) { return a.baseColorTextureIndex < b.baseColorTextureIndex; });
    return model;
}

template<class T>
void appendBytes(std::vector<std::uint8_t>& dst, const T* src, std::size_t count)
{
    const auto* p = reinterpret_cast<const std::uint8_t*>(src);
    dst.insert(dst.end(), p, p + count * sizeof(T));
}
} // namespace
} // namespace rfb200

using namespace rfb200;

// PtFormat::PtFormat(std::filesystem::path gltfPath), pt_format.cpp:20-151.
extern "C" rf_status rf_bake_gltf(const char* gltfPath, rf_pt_file** out)
{
    if (!gltfPath || !out) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_bake_gltf: null argument");
    try
    {
        Model model = loadModel(fs::path(gltfPath));

        // FlattenedModel, flattened_model.cpp:8-46
        std::vector<rf_positions> positions;
        std::vector<float>        normals, texCoords; // 9 / 6 floats per triangle
        std::vector<std::uint32_t> textureIndices;
        for (const Mesh& mesh : model.meshes)
        {
            for (std::size_t i = 0; i < mesh.indices.size(); i += 3)
            {
                rf_positions ps;
                for (int k = 0; k < 3; ++k)
                {
                    const std::uint32_t idx = mesh.indices[i + k];
                    float*              dst = k == 0 ? ps.v0 : k == 1 ? ps.v1 : ps.v2;
                    std::memcpy(dst, &mesh.positions[3 * idx], 12);
                    normals.insert(normals.end(), &mesh.normals[3 * idx], &mesh.normals[3 * idx] + 3);
                    texCoords.insert(texCoords.end(), &mesh.texCoords[2 * idx], &mesh.texCoords[2 * idx] + 2);
                }
                positions.push_back(ps);
                textureIndices.push_back(static_cast<std::uint32_t>(mesh.baseColorTextureIndex));
            }
        }
        if (positions.empty()) return setError(RF_ERROR_FORMAT, "The gltf file %s has no triangles.", gltfPath);

        // buildBvh + reorderAttributes (bvh.hpp:36-46: out[newIndex[i]] = in[i])
        const std::size_t          numTriangles = positions.size();
        std::vector<rf_bvh_node>   nodes(2 * numTriangles);
        std::vector<std::uint64_t> triangleIndices(numTriangles);
        std::uint64_t              numNodes = 0;
        const rf_status            st = rf_build_bvh(positions.data(), numTriangles, nodes.data(), &numNodes, triangleIndices.data());
        if (st != RF_OK) return st;
        nodes.resize(numNodes);

        auto f = std::make_unique<rf_pt_file>();
        std::vector<rf_positions>          bvhPositions(numTriangles);
        std::vector<rf_position_attribute> positionAttributes(numTriangles);
        std::vector<rf_vertex_attributes>  vertexAttributes(numTriangles);
        std::memset(positionAttributes.data(), 0, numTriangles * sizeof(rf_position_attribute)); // padding bytes are zero (pt_format.cpp:65-74)
        std::memset(vertexAttributes.data(), 0, numTriangles * sizeof(rf_vertex_attributes));
        for (std::size_t i = 0; i < numTriangles; ++i)
        {
            const std::size_t dst = triangleIndices[i];
            bvhPositions[dst] = positions[i];
            rf_position_attribute& pa = positionAttributes[dst];
            std::memcpy(pa.p0, positions[i].v0, 12), std::memcpy(pa.p1, positions[i].v1, 12), std::memcpy(pa.p2, positions[i].v2, 12);
            rf_vertex_attributes& va = vertexAttributes[dst];
            std::memcpy(va.n0, &normals[9 * i], 12), std::memcpy(va.n1, &normals[9 * i + 3], 12), std::memcpy(va.n2, &normals[9 * i + 6], 12);
            std::memcpy(va.uv0, &texCoords[6 * i], 8), std::memcpy(va.uv1, &texCoords[6 * i + 2], 8), std::memcpy(va.uv2, &texCoords[6 * i + 4], 8);
            va.texture_idx = textureIndices[i];
        }
        appendBytes(f->arrays[RF_PT_BVH_NODES], nodes.data(), nodes.size());
        appendBytes(f->arrays[RF_PT_BVH_POSITION_ATTRIBUTES], bvhPositions.data(), numTriangles);
        appendBytes(f->arrays[RF_PT_TRIANGLE_POSITION_ATTRIBUTES], positionAttributes.data(), numTriangles);
        appendBytes(f->arrays[RF_PT_TRIANGLE_VERTEX_ATTRIBUTES], vertexAttributes.data(), numTriangles);

        // per-mesh arrays of the rasteriser, pt_format.cpp:84-148
        std::uint64_t vertexOffset = 0, indexOffset = 0;
        for (const Mesh& mesh : model.meshes)
        {
            const std::uint64_t numVertices = mesh.positions.size() / 3, numIndices = mesh.indices.size();
            for (std::uint64_t v = 0; v < numVertices; ++v)
            {
                const float p4[4] = {mesh.positions[3 * v], mesh.positions[3 * v + 1], mesh.positions[3 * v + 2], 1.0f};
                const float n4[4] = {mesh.normals[3 * v], mesh.normals[3 * v + 1], mesh.normals[3 * v + 2], 0.0f};
                appendBytes(f->arrays[RF_PT_VERTEX_POSITIONS], p4, 4);
                appendBytes(f->arrays[RF_PT_VERTEX_NORMALS], n4, 4);
            }
            appendBytes(f->arrays[RF_PT_VERTEX_TEX_COORDS], mesh.texCoords.data(), mesh.texCoords.size());
            appendBytes(f->arrays[RF_PT_VERTEX_INDICES], mesh.indices.data(), mesh.indices.size());
            const std::uint64_t vertexSlice[2] = {vertexOffset, numVertices}, indexSlice[2] = {indexOffset, numIndices};
            appendBytes(f->arrays[RF_PT_MODEL_VERTEX_POSITIONS], vertexSlice, 2);
            appendBytes(f->arrays[RF_PT_MODEL_VERTEX_NORMALS], vertexSlice, 2);
            appendBytes(f->arrays[RF_PT_MODEL_VERTEX_TEX_COORDS], vertexSlice, 2);
            appendBytes(f->arrays[RF_PT_MODEL_VERTEX_INDICES], indexSlice, 2);
            const std::uint32_t textureIndex = static_cast<std::uint32_t>(mesh.baseColorTextureIndex);
            appendBytes(f->arrays[RF_PT_MODEL_BASE_COLOR_TEXTURE_INDICES], &textureIndex, 1);
            vertexOffset += numVertices, indexOffset += numIndices;
        }
        f->textures = std::move(model.textures);
        *out = f.release();
        return RF_OK;
    }
    catch (const std::bad_alloc&)
    {
        return setError(RF_ERROR_IO, "rf_bake_gltf: out of memory");
    }
    catch (const std::exception& e)
    {
        return setError(RF_ERROR_FORMAT, "%s", e.what());
    }
}
