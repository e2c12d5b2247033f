// rm_project.cpp — the data formats either side of the path (host code, setup / delivery, not per sample):
//   * Project::load(path) + Project::build_scene()          core/src/project.rs:13-57
//     serde_json, externally tagged enums:
//       {"objects":[{"geometry":{"Sphere":{"origin":{"x":..,"y":..,"z":..},"radius":..}},
//                    "material":{"Diffuse":[{"x":..,"y":..,"z":..},0.02]}},
//                   {"geometry":{"Plane":{"origin":{..},"normal":{..}}},"material":{"Emission":[{..},{..},0.27,0.0]}},
//                   {"geometry":{"Mesh":"assets/meshes/dragon.ply"},"material":{"Metal":[{..},0.15]}}]}
//     A Mesh becomes Geometry::Grid(Arc::new(AccGrid::build_from_mesh(Mesh::load_ply(path)))) (project.rs:45-49).
//   * the progressive tile message on the wire               server/src/protocol.rs:9-14, core/src/tile.rs:6-14
//       {"type":"TileProgressed","data":{"sample_count":..,"width":..,"height":..,"left":..,"top":..,"data":[{"x":..,"y":..,"z":..},..]}}
//     (serde tag = "type", content = "data"; consumer: editor/src/renderer.js:22-53).

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "rm_internal.hpp"

namespace rm {

int load_ply(const char* path, Mesh* mesh);   // rm_host.cpp

namespace {

struct JValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue* get(const char* key) const {
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char* p;
    const char* end;
    std::string err;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool fail_at(const char* what) { if (err.empty()) err = std::string(what) + " at byte " + std::to_string((long)(p - begin)); return false; }
    const char* begin;
    bool parse(JValue& v, int depth = 0) {
        if (depth > 64) return fail_at("nesting too deep");
        ws();
        if (p >= end) return fail_at("unexpected end of input");
        switch (*p) {
            case '{': {
                v.kind = JValue::Object; p++; ws();
                if (p < end && *p == '}') { p++; return true; }
                for (;;) {
                    ws();
                    JValue k;
                    if (p >= end || *p != '"' || !string(k.str)) return fail_at("expected an object key");
                    ws();
                    if (p >= end || *p != ':') return fail_at("expected ':'");
                    p++;
                    JValue val;
                    if (!parse(val, depth + 1)) return false;
                    v.obj.emplace_back(std::move(k.str), std::move(val));
                    ws();
                    if (p < end && *p == ',') { p++; continue; }
                    if (p < end && *p == '}') { p++; return true; }
                    return fail_at("expected ',' or '}'");
                }
            }
            case '[': {
                v.kind = JValue::Array; p++; ws();
                if (p < end && *p == ']') { p++; return true; }
                for (;;) {
                    JValue e;
                    if (!parse(e, depth + 1)) return false;
                    v.arr.push_back(std::move(e));
                    ws();
                    if (p < end && *p == ',') { p++; continue; }
                    if (p < end && *p == ']') { p++; return true; }
                    return fail_at("expected ',' or ']'");
                }
            }
            case '"': v.kind = JValue::String; return string(v.str);
            case 't': if (end - p >= 4 && !memcmp(p, "true", 4)) { p += 4; v.kind = JValue::Bool; v.b = true; return true; } return fail_at("bad literal");
            case 'f': if (end - p >= 5 && !memcmp(p, "false", 5)) { p += 5; v.kind = JValue::Bool; v.b = false; return true; } return fail_at("bad literal");
            case 'n': if (end - p >= 4 && !memcmp(p, "null", 4)) { p += 4; v.kind = JValue::Null; return true; } return fail_at("bad literal");
            default: {
                const char* q = p;
                if (q < end && *q == '-') q++;
                if (q >= end || !(*q >= '0' && *q <= '9')) return fail_at("unexpected character");
                auto r = std::from_chars(p, end, v.num);
                if (r.ec != std::errc()) return fail_at("bad number");
                p = r.ptr;
                v.kind = JValue::Number;
                return true;
            }
        }
    }
    bool string(std::string& out) {
        p++;   // opening quote
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail_at("bad escape");
                switch (*p) {
                    case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
                    case 'r': out += '\r'; break; case 't': out += '\t'; break;
                    case 'u': {
                        if (end - p < 5) return fail_at("bad \\u escape");
                        unsigned cp = 0;
                        for (int i = 1; i <= 4; i++) {
                            const char c = p[i];
                            cp = cp * 16 + (c >= '0' && c <= '9' ? (unsigned)(c - '0') : c >= 'a' && c <= 'f' ? (unsigned)(c - 'a' + 10) : c >= 'A' && c <= 'F' ? (unsigned)(c - 'A' + 10) : 0x10000u);
                        }
                        if (cp > 0xffff) return fail_at("bad \\u escape");
                        p += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xc0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3f)); }
                        else { out += (char)(0xe0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3f)); out += (char)(0x80 | (cp & 0x3f)); }
                        break;
                    }
                    default: return fail_at("bad escape");
                }
                p++;
            } else {
                out += *p++;
            }
        }
        if (p >= end) return fail_at("unterminated string");
        p++;
        return true;
    }
};

bool as_f64(const JValue* v, double* out) { if (!v || v->kind != JValue::Number) return false; *out = v->num; return true; }
bool as_vec3(const JValue* v, rm_vec3* out) {
    return v && v->kind == JValue::Object && as_f64(v->get("x"), &out->x) && as_f64(v->get("y"), &out->y) && as_f64(v->get("z"), &out->z);
}
// an externally tagged enum value: {"Variant": content}
bool variant(const JValue* v, std::string* name, const JValue** content) {
    if (!v || v->kind != JValue::Object || v->obj.size() != 1) return false;
    *name = v->obj[0].first;
    *content = &v->obj[0].second;
    return true;
}

}  // namespace

int load_project(const char* path, rm_scene* scene) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return fail(RM_ERR_IO, std::string("cannot open ") + path);               // File::open(p)?
    const std::streamsize len = f.tellg();
    f.seekg(0);
    std::string buf((size_t)std::max<std::streamsize>(len, 0), '\0');
    if (len > 0 && !f.read(&buf[0], len)) return fail(RM_ERR_IO, std::string("cannot read ") + path);
    JParser jp{buf.data(), buf.data() + buf.size(), std::string(), buf.data()};
    JValue root;
    if (!jp.parse(root)) return fail(RM_ERR_PROJECT, "project JSON: " + jp.err);
    jp.ws();
    if (jp.p != jp.end) return fail(RM_ERR_PROJECT, "project JSON: trailing characters");
    const JValue* objects = root.kind == JValue::Object ? root.get("objects") : nullptr;
    if (!objects || objects->kind != JValue::Array) return fail(RM_ERR_PROJECT, "project JSON: missing field `objects`");
    for (size_t i = 0; i < objects->arr.size(); i++) {
        const JValue& o = objects->arr[i];
        const std::string where = "project JSON: objects[" + std::to_string(i) + "]";
        std::string gname, mname;
        const JValue *gc = nullptr, *mc = nullptr;
        if (o.kind != JValue::Object || !variant(o.get("geometry"), &gname, &gc) || !variant(o.get("material"), &mname, &mc))
            return fail(RM_ERR_PROJECT, where + ": expected {\"geometry\": {Variant: ..}, \"material\": {Variant: ..}}");
        Object ob{};
        // enum Material { Diffuse(Vector3, f64), Metal(Vector3, f64), Emission(Vector3, Vector3, f64, f64) }   core/src/lib.rs:21-26
        rm_material& m = ob.material;
        if ((mname == "Diffuse" || mname == "Metal") && mc->kind == JValue::Array && mc->arr.size() == 2 && as_vec3(&mc->arr[0], &m.a) && as_f64(&mc->arr[1], &m.p0)) {
            m.kind = mname == "Diffuse" ? RM_MATERIAL_DIFFUSE : RM_MATERIAL_METAL;
        } else if (mname == "Emission" && mc->kind == JValue::Array && mc->arr.size() == 4 && as_vec3(&mc->arr[0], &m.a) && as_vec3(&mc->arr[1], &m.b) &&
                   as_f64(&mc->arr[2], &m.p0) && as_f64(&mc->arr[3], &m.p1)) {
            m.kind = RM_MATERIAL_EMISSION;
        } else {
            return fail(RM_ERR_PROJECT, where + ": bad material `" + mname + "`");
        }
        if (gname == "Sphere" && gc->kind == JValue::Object && as_vec3(gc->get("origin"), &ob.origin) && as_f64(gc->get("radius"), &ob.radius)) {
            ob.geometry = GEOM_SPHERE;
        } else if (gname == "Plane" && gc->kind == JValue::Object && as_vec3(gc->get("origin"), &ob.origin) && as_vec3(gc->get("normal"), &ob.normal)) {
            ob.geometry = GEOM_PLANE;
        } else if (gname == "Mesh" && gc->kind == JValue::String) {
            Mesh mesh;
            if (int st = load_ply(gc->str.c_str(), &mesh)) return st;                  // Mesh::load_ply(m)
            rm_aabb bounds = mesh.bounds;
            if (int st = build_grid(std::move(mesh.triangles), bounds, &ob.grid)) return st;   // AccGrid::build_from_mesh
            ob.geometry = GEOM_GRID;
        } else {
            return fail(RM_ERR_PROJECT, where + ": bad geometry `" + gname + "`");
        }
        scene->objects.push_back(std::move(ob));
    }
    return RM_OK;
}

// ------------------------------------------------------------------ tile message -> JSON

namespace {
void put_f64(std::string& out, double v) {
    if (!std::isfinite(v)) { out += "null"; return; }          // serde_json writes non-finite f64 as null
    char tmp[32];
    auto r = std::to_chars(tmp, tmp + sizeof(tmp), v);          // shortest representation that round-trips
    out.append(tmp, r.ptr);
    // serde_json always marks a float: 1.0, not 1
    bool marked = false;
    for (const char* c = tmp; c < r.ptr; c++) if (*c == '.' || *c == 'e' || *c == 'n' || *c == 'i') marked = true;
    if (!marked) out += ".0";
}
void put_usize(std::string& out, size_t v) { out += std::to_string(v); }
}  // namespace

std::string message_json(const rm_message& m) {
    std::string out;
    const rm_tile& t = m.tile;
    out.reserve(128 + t.width * t.height * 72);
    out += "{\"type\":\"";
    out += m.kind == RM_TILE_FINISHED ? "TileFinished" : "TileProgressed";
    out += "\",\"data\":{\"sample_count\":"; put_usize(out, t.sample_count);
    out += ",\"width\":"; put_usize(out, t.width);
    out += ",\"height\":"; put_usize(out, t.height);
    out += ",\"left\":"; put_usize(out, t.left);
    out += ",\"top\":"; put_usize(out, t.top);
    out += ",\"data\":[";
    for (size_t i = 0; i < t.width * t.height; i++) {
        if (i) out += ',';
        out += "{\"x\":"; put_f64(out, t.data[i].x);
        out += ",\"y\":"; put_f64(out, t.data[i].y);
        out += ",\"z\":"; put_f64(out, t.data[i].z);
        out += '}';
    }
    out += "]}}";
    return out;
}

}  // namespace rm

using namespace rm;

extern "C" {

rm_scene* rm_project_load_scene(const char* path, int* status) {
    int st = RM_OK;
    rm_scene* s = nullptr;
    if (!path) {
        st = fail(RM_ERR_INVALID_ARGUMENT, "rm_project_load_scene: null path");
    } else {
        try {
            s = new rm_scene();
            st = load_project(path, s);
        } catch (const std::bad_alloc&) {
            st = fail(RM_ERR_OUT_OF_MEMORY, "out of memory while loading the project");
        }
        if (st != RM_OK) { delete s; s = nullptr; }
    }
    if (status) *status = st;
    return s;
}

size_t rm_message_to_json(const rm_message* message, char* buffer, size_t capacity) {
    if (!message || (!message->tile.data && message->tile.width != 0 && message->tile.height != 0)) return 0;
    try {
        const std::string s = message_json(*message);
        if (buffer && capacity) {
            const size_t n = std::min(s.size(), capacity - 1);
            memcpy(buffer, s.data(), n);
            buffer[n] = '\0';
        }
        return s.size();
    } catch (const std::bad_alloc&) {
        return 0;
    }
}

}  // extern "C"
