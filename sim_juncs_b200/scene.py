"""Flattened scene: what bound_geom extracts from a parsed .geom file (reference src/disp.cpp).

A scene arrives as the JSON document written by a CGS front end -- today the reference's own parser
through oracle/_ref/scene_dump (test-time) or the fixtures committed under scenes/*.json; the JSON
schema is documented in DESIGN.md.  This module turns it into:
  * regions   : flattened CSG node arrays + eps + poles            (disp.cpp:482-548)
  * sources   : source_info records                                (disp.cpp:318-376)
  * monitors  : locations + cluster boundaries                     (disp.cpp:456-475, 627-637)
  * cgs_params: every numeric context variable                     (disp.cpp:849-862)
"""
import json
import math

from ._lib import SjCsgNode

LIGHT_SPEED = 0.299792458
THICK_SCALE = 1.0
_TYPES = {"composite": 0, "sphere": 1, "box": 2, "plane": 3, "cylinder": 4, "undef": 5}
COMPONENTS = ["Ex", "Ey", "Ez", "Hx", "Hy", "Hz"]


def _num(x):
    if isinstance(x, str):
        return {"nan": math.nan, "inf": math.inf, "-inf": -math.inf}[x]
    return float(x)


class Inst:
    """A VAL_INST value: ordered (name, value) pairs; lookup returns the newest match."""

    def __init__(self, fields):
        self.fields = fields

    def lookup(self, name):
        for k, v in reversed(self.fields):
            if k == name:
                return v
        return None

    @property
    def type(self):
        t = self.lookup("__type__")
        return t if isinstance(t, str) else None


def _decode(v):
    if isinstance(v, dict):
        if "__fields__" in v:
            return Inst([(k, _decode(x)) for k, x in v["__fields__"]])
        if "vec3" in v:
            return ("vec3", [_num(x) for x in v["vec3"]])
        if "mat" in v:
            return ("mat", [_num(x) for x in v["mat"]])
        if "func" in v:
            return ("func",)
        return v
    if isinstance(v, list):
        return [_decode(x) for x in v]
    if isinstance(v, str) and v in ("nan", "inf", "-inf"):
        return v
    return v


def _as_vec3(v):
    """value::cast_to(VAL_3VEC): vec3 stays, a 3-element numeric list converts."""
    if isinstance(v, tuple) and v and v[0] == "vec3":
        return list(v[1])
    if isinstance(v, list) and len(v) == 3 and all(isinstance(x, (int, float)) for x in v):
        return [float(x) for x in v]
    return None


def flatten_tree(tree, nodes):
    """Depth-first flattening of a dumped CSG tree into SjCsgNode records; returns the node index."""
    if tree is None:
        return -1
    idx = len(nodes)
    nd = SjCsgNode()
    nodes.append(nd)
    nd.type = _TYPES.get(tree["type"], 5)
    nd.invert = int(tree.get("invert", 0))
    nd.cmb = int(tree.get("cmb", 0))
    nd.child0 = nd.child1 = -1
    M = tree.get("M", [1, 0, 0, 0, 1, 0, 0, 0, 1])
    for i in range(9):
        nd.M[i] = _num(M[i])
    t = tree["type"]
    if t == "sphere":
        p = list(tree["center"]) + [tree["rad"]]
    elif t == "box":
        p = list(tree["center"]) + list(tree["offset"])
    elif t == "plane":
        p = list(tree["normal"]) + [tree["offset"]]
    elif t == "cylinder":
        p = list(tree["center"]) + [tree["height"], tree["r1_sq"], tree["r1_sq_x_h"], tree["r2_sq"]]
    else:
        p = []
    for i, x in enumerate(p):
        nd.p[i] = _num(x)
    if t == "composite":
        ch = tree.get("children", [None, None])
        nd.child0 = flatten_tree(ch[0], nodes)
        nd.child1 = flatten_tree(ch[1], nodes)
    return idx


def parse_susceptibilities(val):
    """bound_geom::parse_susceptibilities (disp.cpp:418-454).  Returns (poles, ercode); poles are
    (omega_0, gamma, sigma, use_denom).  On a malformed entry it returns what was read so far with
    -1/-2 -- the caller ignores the code (disp.cpp:534), which is how Au loses its Lorentz pole."""
    ret = []
    if isinstance(val, list):
        for cur in val:
            if not isinstance(cur, list) or len(cur) < 2:
                return ret, -1
            # the reference reads l[2] without a length check; a 2-element list is UB there.  Treat
            # it as the -2 "bad argument" exit, which is what the shipped Au entry produces [probe].
            if len(cur) < 3 or not all(isinstance(cur[i], (int, float)) for i in range(3)):
                return ret, -2
            use_denom = True
            if len(cur) > 3:
                if isinstance(cur[3], (int, float)):
                    use_denom = (cur[3] != 0.0)
                elif isinstance(cur[3], str):
                    tok = cur[3]
                    if tok in ("drude", "true"):
                        use_denom = False
                    if tok in ("lorentz", "false"):
                        use_denom = True
                else:
                    return ret, -2
            ret.append((float(cur[0]), float(cur[1]), float(cur[2]), use_denom))
    return ret, 0


class SourceInfo:
    """source_info (disp.cpp:318-376): raw .geom units (um, fs)."""

    def __init__(self, inst):
        self.ok = False
        self.type = "gaussian"
        self.component = 0
        self.wavelen = 0.7
        self.width = 1.0
        self.phase = 0.0
        self.start_time = 0.0
        self.end_time = 0.0
        self.amplitude = 1.0
        if not isinstance(inst, Inst):
            return
        is_gauss = inst.type == "Gaussian_source"
        is_contin = inst.type == "CW_source"
        if not (is_gauss or is_contin):
            return
        self.ok = True

        def num(name):
            v = inst.lookup(name)
            return float(v) if isinstance(v, (int, float)) else None
        c = num("component")
        if c is not None and int(c) in (1, 2, 3, 4, 5):
            self.component = int(c)
        for key, attr in (("wavelength", "wavelen"), ("amplitude", "amplitude"), ("start_time", "start_time"),
                          ("end_time", "end_time"), ("slowness", "width")):
            v = num(key)
            if v is not None:
                setattr(self, attr, v)
        if is_gauss:
            v = num("width")
            if v is not None:
                self.width = v
            v = num("phase")
            if v is not None:
                self.phase = v
            cutoff = 5
            if num("cutoff") is not None:
                cutoff = 6          # sic: any numeric cutoff becomes 6 (disp.cpp:369-372)
            self.end_time = self.start_time + 2 * cutoff * self.width
        if is_contin:
            self.type = "continuous"
        self.region = inst.lookup("region")


class Region:
    def __init__(self):
        self.root = -1
        self.eps = None
        self.poles_raw = []      # (omega0, gamma, sigma, use_denom) in file units
        self.sus_ercode = 0
        self.make_2d = False
        self.thickness = 1.0
        self.metadata = {}


class Scene:
    def __init__(self, doc):
        self.ercode = int(doc.get("ercode", 0))
        self.nodes = []
        self.regions = []
        for r in doc.get("roots", []):
            reg = Region()
            reg.metadata = {k: _decode(v) for k, v in r.get("metadata", {}).items()}
            reg.root = flatten_tree(r["tree"], self.nodes)
            eps = reg.metadata.get("eps")
            reg.eps = float(eps) if isinstance(eps, (int, float)) else None
            m2 = reg.metadata.get("make_2d")
            reg.make_2d = isinstance(m2, (int, float)) and m2 != 0
            reg.thickness = _num(r.get("thickness", 1.0))
            if "susceptibilities" in reg.metadata:
                reg.poles_raw, reg.sus_ercode = parse_susceptibilities(reg.metadata["susceptibilities"])
            self.regions.append(reg)
        self.context = _decode(doc.get("context", {"__fields__": []}))
        # bound_geom ctor, disp.cpp:584-639: walk the context oldest -> newest
        self.sources = []
        self.source_boxes = []
        self.monitor_locs = []
        self.monitor_clusters = []
        for name, inst in self.context.fields:
            info = SourceInfo(inst)
            if info.ok:
                reg = info.region
                if isinstance(reg, Inst) and reg.type == "Box":
                    p1, p2 = _as_vec3(reg.lookup("pt_1")), _as_vec3(reg.lookup("pt_2"))
                    if p1 is not None and p2 is not None:
                        self.sources.append(info)
                        self.source_boxes.append((p1, p2))
            elif isinstance(inst, Inst) and inst.type == "monitor":
                vl = inst.lookup("locations")
                if isinstance(vl, list):
                    for v in vl:
                        p = _as_vec3(v)
                        if p is None:
                            break
                        self.monitor_locs.append(p)
                self.monitor_clusters.append(len(self.monitor_locs))
        self.cgs_params = [(k, float(v)) for k, v in self.context.fields
                           if isinstance(v, (int, float)) and not isinstance(v, bool) and k]

    def node_array(self):
        arr = (SjCsgNode * max(len(self.nodes), 1))()
        for i, nd in enumerate(self.nodes):
            arr[i] = nd
        return arr

    @staticmethod
    def load(path):
        with open(path, "r") as fp:
            return Scene(json.load(fp))

    @staticmethod
    def from_geom(path, settings):
        """read a .geom file with this repo's own CGS reader (sim_juncs_b200/cgs.py); the scene::scene(fname,
        context) call of bound_geom's constructor, disp.cpp:556."""
        from . import cgs
        return Scene(cgs.parse_geom(path, settings))
