// sj_flatten.cpp -- the one place that needs the CSG classes' protected members (trans_mat, invert,
// centres ...): the reference exposes no getters for them, so this translation unit opens the
// access specifiers before including the reference header.  A maintainer would add getters instead.
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <cstdint>
#include <cmath>
#define private public
#define protected public
#include "cgs.hpp"
#undef private
#undef protected
#include "sim_juncs_b200.h"

static int put_node(object *o, object_type t, std::vector<sj_csg_node> &nodes) {
    if (!o) return -1;
    const int idx = (int)nodes.size();
    sj_csg_node nd;
    memset(&nd, 0, sizeof nd);
    nd.child0 = nd.child1 = -1;
    nd.type = 5;
    nodes.push_back(nd);
    for (int i = 0; i < 9; ++i) nd.M[i] = o->trans_mat.el[i];
    nd.invert = o->invert;
    switch (t) {
        case CGS_SPHERE: { sphere *s = (sphere *)o; nd.type = 1; for (int i = 0; i < 3; ++i) nd.p[i] = s->center.el[i]; nd.p[3] = s->rad; break; }
        case CGS_BOX: { box *b = (box *)o; nd.type = 2; for (int i = 0; i < 3; ++i) { nd.p[i] = b->center.el[i]; nd.p[3 + i] = b->offset.el[i]; } break; }
        case CGS_PLANE: { plane *q = (plane *)o; nd.type = 3; for (int i = 0; i < 3; ++i) nd.p[i] = q->normal.el[i]; nd.p[3] = q->offset; break; }
        case CGS_CYLINDER: { cylinder *c = (cylinder *)o; nd.type = 4; for (int i = 0; i < 3; ++i) nd.p[i] = c->center.el[i];
                             nd.p[3] = c->height; nd.p[4] = c->r1_sq; nd.p[5] = c->r1_sq_x_h; nd.p[6] = c->r2_sq; break; }
        case CGS_COMPOSITE: case CGS_ROOT: case CGS_DATA: {
            composite_object *c = (composite_object *)o;
            nd.type = 0; nd.cmb = (int)c->cmb;
            const int l = put_node(c->children[0], c->child_types[0], nodes);
            const int r = put_node(c->children[1], c->child_types[1], nodes);
            nd.child0 = l; nd.child1 = r;
            break; }
        default: break;
    }
    nodes[idx] = nd;
    return idx;
}

int sj_flatten_tree(composite_object *root, std::vector<sj_csg_node> &nodes) { return put_node(root, CGS_COMPOSITE, nodes); }
