/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE.  A small driver (our code) that is compiled together
 * with the reference's own scene-language sources IN PLACE from /root/reference/src
 * (cgs.cpp, cgs_data.cpp, cgs_read.cpp; see oracle/Makefile target `ref`) into
 * oracle/_ref/scene_dump.  No reference source is copied into this repository.
 *
 * It is the strongest oracle available for the voxelization half of the hot path
 * (SURVEY.md section 8c): it runs the reference parser + composite_object::in()
 * (reference src/cgs.cpp:403-447) unmodified and
 *   dump   : writes the parsed scene as JSON (flattened CSG trees with every private field,
 *            metadata, the full evaluation context incl. sources and monitors);
 *   points : reads N xyz triples (float64, binary) and writes one byte per point whose bit r is
 *            roots[r]->in(point) -- the golden inside masks.
 *
 * The evaluation context is seeded exactly like the reference driver does it
 * (reference src/disp.cpp:28-51, context_from_settings) -- that file itself cannot be compiled
 * here because it includes <meep.hpp>, so the dozen lines that seed the context are restated below.
 *
 * usage: scene_dump dump   <geom> <pml_thickness> <sim_length> <um_scale> <out_dir> <opts|-> <resolution>
 *        scene_dump points <geom> <pml_thickness> <sim_length> <um_scale> <out_dir> <opts|-> <resolution> <in.f64> <out.u8>
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unordered_map>
#include <cstdint>
#include <cmath>
#include <unistd.h>

#define private public
#define protected public
#include "cgs.hpp"
#undef private
#undef protected

#define LIGHT_SPEED 0.299792458
#define THICK_SCALE 1.0

static value um_to_l(context& c, cgs_func f, parse_ercode& er) {
    value ret;
    if (f.n_args < 1) { er = E_LACK_TOKENS; return ret; }
    if (f.args[0].type != VAL_NUM) { er = E_BAD_TOKEN; return ret; }
    value l_per_um = c.lookup("l_per_um");
    if (l_per_um.type != VAL_NUM) { er = E_NOT_DEFINED; return ret; }
    return make_val_num((l_per_um.val.x) * (f.args[0].val.x));
}
static value fs_to_t(context& c, cgs_func f, parse_ercode& er) {
    value ret;
    if (f.n_args < 1) { er = E_LACK_TOKENS; return ret; }
    if (f.args[0].type != VAL_NUM) { er = E_BAD_TOKEN; return ret; }
    value l_per_um = c.lookup("l_per_um");
    if (l_per_um.type != VAL_NUM) { er = E_NOT_DEFINED; return ret; }
    return make_val_num(LIGHT_SPEED * (l_per_um.val.x) * (f.args[0].val.x));
}

static context make_context(double pml, double len, double um_scale, const char* out_dir, const char* opts) {
    context con;
    con.emplace("pi", make_val_num(M_PI));
    con.emplace("pml_thickness", make_val_num(pml));
    con.emplace("sim_length", make_val_num(len));
    con.emplace("length", make_val_num(2 * pml + len));
    con.emplace("l_per_um", make_val_num(um_scale));
    value tmp_out = make_val_str(out_dir);
    con.emplace("out_dir", tmp_out);
    cleanup_val(&tmp_out);
    value tmp_f = make_val_func("um_to_l", 1, &um_to_l);
    con.emplace("um_to_l", tmp_f);
    cleanup_val(&tmp_f);
    tmp_f = make_val_func("fs_to_t", 1, &fs_to_t);
    con.emplace("fs_to_t", tmp_f);
    cleanup_val(&tmp_f);
    if (opts && strcmp(opts, "-") != 0) {
        line_buffer lb(opts, ';');
        con.read_from_lines(lb);
    }
    return con;
}

static void jstr(FILE* fp, const char* s) {
    fputc('"', fp);
    for (; s && *s; ++s) {
        if (*s == '"' || *s == '\\') fputc('\\', fp);
        if (*s == '\n') { fputs("\\n", fp); continue; }
        fputc(*s, fp);
    }
    fputc('"', fp);
}
static void jnum(FILE* fp, double x) {
    if (std::isnan(x)) fputs("\"nan\"", fp);
    else if (std::isinf(x)) fputs(x > 0 ? "\"inf\"" : "\"-inf\"", fp);
    else fprintf(fp, "%.17g", x);
}

static void jvalue(FILE* fp, value v, int depth);
static void jcontext(FILE* fp, context* c, int depth) {
    fputs("{\"__fields__\":[", fp);
    size_t n = c->size();
    for (size_t i = n; i > 0; --i) {  // oldest first
        name_val_pair p = c->peek(i);
        if (i != n) fputc(',', fp);
        fputc('[', fp);
        jstr(fp, p.get_name() ? p.get_name() : "");
        fputc(',', fp);
        jvalue(fp, p.get_val(), depth + 1);
        fputc(']', fp);
    }
    fputs("]}", fp);
}
static void jvalue(FILE* fp, value v, int depth) {
    if (depth > 12) { fputs("null", fp); return; }
    switch (v.type) {
        case VAL_NUM: jnum(fp, v.val.x); break;
        case VAL_STR: jstr(fp, v.val.s); break;
        case VAL_3VEC:
            fputs("{\"vec3\":[", fp); jnum(fp, v.val.v->el[0]); fputc(',', fp); jnum(fp, v.val.v->el[1]);
            fputc(',', fp); jnum(fp, v.val.v->el[2]); fputs("]}", fp); break;
        case VAL_LIST:
            fputc('[', fp);
            for (size_t i = 0; i < v.n_els; ++i) { if (i) fputc(',', fp); jvalue(fp, v.val.l[i], depth + 1); }
            fputc(']', fp); break;
        case VAL_INST: jcontext(fp, v.val.c, depth); break;
        case VAL_FUNC: fputs("{\"func\":true}", fp); break;
        case VAL_MAT:
            fputs("{\"mat\":[", fp);
            for (int i = 0; i < 9; ++i) { if (i) fputc(',', fp); jnum(fp, v.val.m->el[i]); }
            fputs("]}", fp); break;
        default: fputs("null", fp);
    }
}

static void jmat(FILE* fp, const mat3x3& m) {
    fputs("\"M\":[", fp);
    for (int i = 0; i < 9; ++i) { if (i) fputc(',', fp); jnum(fp, m.el[i]); }
    fputc(']', fp);
}
static void jv3(FILE* fp, const char* name, const vec3& v) {
    fprintf(fp, "\"%s\":[", name); jnum(fp, v.el[0]); fputc(',', fp); jnum(fp, v.el[1]); fputc(',', fp); jnum(fp, v.el[2]); fputc(']', fp);
}

static void jnode(FILE* fp, object* o, object_type t) {
    if (!o) { fputs("null", fp); return; }
    switch (t) {
        case CGS_SPHERE: {
            sphere* s = (sphere*)o;
            fprintf(fp, "{\"type\":\"sphere\",\"invert\":%d,", s->invert); jmat(fp, s->trans_mat); fputc(',', fp);
            jv3(fp, "center", s->center); fputs(",\"rad\":", fp); jnum(fp, s->rad); fputc('}', fp); break; }
        case CGS_BOX: {
            box* b = (box*)o;
            fprintf(fp, "{\"type\":\"box\",\"invert\":%d,", b->invert); jmat(fp, b->trans_mat); fputc(',', fp);
            jv3(fp, "center", b->center); fputc(',', fp); jv3(fp, "offset", b->offset); fputc('}', fp); break; }
        case CGS_PLANE: {
            plane* p = (plane*)o;
            fprintf(fp, "{\"type\":\"plane\",\"invert\":%d,", p->invert); jmat(fp, p->trans_mat); fputc(',', fp);
            jv3(fp, "normal", p->normal); fputs(",\"offset\":", fp); jnum(fp, p->offset); fputc('}', fp); break; }
        case CGS_CYLINDER: {
            cylinder* c = (cylinder*)o;
            fprintf(fp, "{\"type\":\"cylinder\",\"invert\":%d,", c->invert); jmat(fp, c->trans_mat); fputc(',', fp);
            jv3(fp, "center", c->center); fputs(",\"height\":", fp); jnum(fp, c->height);
            fputs(",\"r1_sq\":", fp); jnum(fp, c->r1_sq); fputs(",\"r1_sq_x_h\":", fp); jnum(fp, c->r1_sq_x_h);
            fputs(",\"r2_sq\":", fp); jnum(fp, c->r2_sq); fputc('}', fp); break; }
        case CGS_COMPOSITE: case CGS_ROOT: case CGS_DATA: {
            composite_object* c = (composite_object*)o;
            fprintf(fp, "{\"type\":\"composite\",\"invert\":%d,\"cmb\":%d,", c->invert, (int)c->cmb); jmat(fp, c->trans_mat);
            fprintf(fp, ",\"child_types\":[%d,%d],\"children\":[", (int)c->child_types[0], (int)c->child_types[1]);
            jnode(fp, c->children[0], c->child_types[0]); fputc(',', fp);
            jnode(fp, c->children[1], c->child_types[1]);
            fputs("]}", fp); break; }
        default:
            fprintf(fp, "{\"type\":\"undef\",\"code\":%d}", (int)t);
    }
}

int main(int argc, char** argv) {
    if (argc < 9) {
        fprintf(stderr, "usage: %s dump|points <geom> <pml> <len> <um_scale> <out_dir> <opts|-> <resolution> [in.f64 out.u8]\n", argv[0]);
        return 2;
    }
    const char* mode = argv[1];
    const char* geom = argv[2];
    double pml = atof(argv[3]), len = atof(argv[4]), um = atof(argv[5]);
    const char* out_dir = argv[6];
    const char* opts = argv[7];
    double resolution = atof(argv[8]);

    // the reference prints from print()/snapshot while parsing: keep stdout for our payload only
    FILE* payload = fdopen(dup(fileno(stdout)), "w");
    if (!freopen("/dev/null", "w", stdout)) return 3;

    parse_ercode er = E_SUCCESS;
    scene problem(geom, make_context(pml, len, um, out_dir, opts), &er);
    std::vector<composite_object*> roots = problem.get_roots();

    // reference src/disp.cpp:518-525: make_2d roots are rescaled along z before any inside test.
    // Only the `points` mode bakes this in (it needs the resolution); `dump` reports the raw tree
    // plus the make_2d flag and leaves the rescale to the consumer.
    std::vector<double> thick(roots.size(), 1.0);
    if (strcmp(mode, "points") == 0)
        for (size_t i = 0; i < roots.size(); ++i) {
            if (roots[i]->has_metadata("make_2d") && roots[i]->fetch_metadata("make_2d").val.x != 0) {
                thick[i] = THICK_SCALE / resolution;
                roots[i]->rescale(vec3(1.0, 1.0, thick[i]));
            }
        }

    if (strcmp(mode, "dump") == 0) {
        FILE* fp = payload;
        fprintf(fp, "{\"ercode\":%d,\"n_roots\":%zu,\"roots\":[", (int)er, roots.size());
        for (size_t i = 0; i < roots.size(); ++i) {
            if (i) fputc(',', fp);
            fputs("{\"thickness\":", fp); jnum(fp, thick[i]);
            fputs(",\"metadata\":{", fp);
            bool first = true;
            for (auto it = roots[i]->metadata.begin(); it != roots[i]->metadata.end(); ++it) {
                if (!first) fputc(',', fp);
                first = false;
                jstr(fp, it->first.c_str()); fputc(':', fp); jvalue(fp, it->second, 0);
            }
            fputs("},\"tree\":", fp);
            jnode(fp, roots[i], CGS_COMPOSITE);
            fputc('}', fp);
        }
        fputs("],\"context\":", fp);
        jcontext(fp, &problem.get_context(), 0);
        fputs("}\n", fp);
        fclose(fp);
        return 0;
    }
    if (strcmp(mode, "points") == 0 && argc >= 11) {
        FILE* fi = fopen(argv[9], "rb");
        FILE* fo = fopen(argv[10], "wb");
        if (!fi || !fo) return 4;
        std::vector<double> buf(3 * 65536);
        std::vector<uint8_t> out(65536);
        size_t got;
        while ((got = fread(buf.data(), 3 * sizeof(double), 65536, fi)) > 0) {
            for (size_t p = 0; p < got; ++p) {
                uint8_t m = 0;
                vec3 r(buf[3 * p], buf[3 * p + 1], buf[3 * p + 2]);
                for (size_t i = 0; i < roots.size() && i < 8; ++i)
                    if (roots[i]->in(r)) m |= (uint8_t)(1u << i);
                out[p] = m;
            }
            fwrite(out.data(), 1, got, fo);
        }
        fclose(fi); fclose(fo);
        fprintf(payload, "{\"ercode\":%d,\"n_roots\":%zu}\n", (int)er, roots.size());
        fclose(payload);
        return 0;
    }
    return 2;
}
