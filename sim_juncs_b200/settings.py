"""parse_settings mirror: params.conf + CLI flags with the reference's exact semantics
(reference src/argparse.h:27-47 struct, :120-169 handle_pair, :171-191 correct_defaults,
:196-241 parse_conf_file, :244-383 parse_args)."""
DEFAULT_PML_THICK = 0.3
DEFAULT_LEN = 8.0
DEFAULT_RESOLUTION = 2.0
DEFAULT_OUT_DIR = "/tmp"
DEFAULT_SMOOTH_RAD = 0.05


def _strtod(s):
    """C strtod(): longest valid numeric prefix, 0.0 when there is none."""
    s = s.lstrip(" \t\n")
    best = 0.0
    for end in range(len(s), 0, -1):
        try:
            best = float(s[:end])
            return best
        except ValueError:
            continue
    return best


def _strtol(s):
    s = s.lstrip(" \t\n")
    i = 0
    if i < len(s) and s[i] in "+-":
        i += 1
    while i < len(s) and s[i].isdigit():
        i += 1
    try:
        return int(s[:i])
    except ValueError:
        return 0


class ParseSettings:
    """Field-for-field mirror of the reference's parse_settings (argparse.h:27-47)."""

    def __init__(self):
        self.out_dir = None
        self.geom_fname = None
        self.conf_fname = None
        self.n_dims = -3
        self.pml_thickness = -DEFAULT_PML_THICK
        self.len = -DEFAULT_LEN
        self.um_scale = 1.0
        self.grid_num = -1
        self.resolution = -DEFAULT_RESOLUTION
        self.courant = 0.5
        self.ambient_eps = -1.0
        self.smooth_n = 0
        self.smooth_rad = DEFAULT_SMOOTH_RAD
        self.post_source_t = 10.0
        self.save_span = 20
        self.dump_raw = 0
        self.verbosity = 1
        self.user_opts = None

    # argparse.h:120-169.  Keys guarded by a negative / NULL sentinel do not override a CLI value.
    def handle_pair(self, tok, val):
        tok = tok.rstrip(" \t")
        val = val.rstrip(" \t")
        if tok == "pml_thickness" and self.pml_thickness < 0:
            self.pml_thickness = _strtod(val)
        elif tok == "resolution" and self.resolution < 0:
            self.resolution = _strtod(val)
        elif tok == "dimensions" and self.n_dims < 0:
            self.n_dims = _strtol(val)
        elif tok == "um_scale":
            self.um_scale = _strtod(val)
        elif tok == "length" and self.len < 0:
            self.len = _strtod(val)
        elif tok == "smooth_rad":
            self.smooth_rad = _strtod(val)
        elif tok == "smooth_n":
            self.smooth_n = _strtol(val)
        elif tok == "courant":
            self.courant = _strtod(val)
        elif tok == "ambient_eps" and self.ambient_eps < 0:
            self.ambient_eps = _strtod(val)
        elif tok == "geom_fname" and not self.geom_fname:
            self.geom_fname = val.strip(" \t\n")
        elif tok == "out_dir" and not self.out_dir:
            self.out_dir = val.strip(" \t\n")
        elif tok == "post_source_t":
            self.post_source_t = _strtod(val)
        elif tok == "save_span":
            self.save_span = _strtol(val)
        elif tok == "dump_raw":
            if val == "true":
                self.dump_raw = 1
            elif val == "false":
                self.dump_raw = 0
            else:
                self.dump_raw = _strtol(val)

    # argparse.h:171-191
    def correct_defaults(self):
        total_len = self.len + 2 * self.pml_thickness
        if self.grid_num < 0:
            # ensure that we have an odd number of grid points
            self.grid_num = 2 * int(0.5 * (1 + total_len * self.resolution)) + 1
        self.n_dims = abs(self.n_dims)
        self.pml_thickness = abs(self.pml_thickness)
        self.len = abs(self.len)
        self.resolution = abs(self.resolution)
        self.ambient_eps = abs(self.ambient_eps)
        self.resolution = float(self.grid_num) / total_len
        if self.out_dir is None:
            self.out_dir = DEFAULT_OUT_DIR

    # argparse.h:196-241: `[section]` prefixes skipped, `#` comments, `key = value`, `;` separates
    def parse_conf_file(self, fname):
        with open(fname, "r") as fp:
            for raw in fp:
                buf = raw
                i = 0
                while i < len(buf) and buf[i] in " \t":
                    i += 1
                if i < len(buf) and buf[i] == "[":
                    while i < len(buf) and buf[i] != "]":
                        i += 1
                    i += 1
                tok_start = i
                tok = None
                val_start = None
                while i < len(buf) and buf[i] != "#":
                    ch = buf[i]
                    if val_start is None:
                        if ch == "=":
                            tok = buf[tok_start:i]
                            val_start = i + 1
                    elif ch == "\n" or ch == ";":
                        self.handle_pair(tok, buf[val_start:i])
                        tok_start = i + 1
                        val_start = None
                    i += 1
        self.correct_defaults()
        return 0

    # argparse.h:244-383 (flags are matched by prefix, as strstr(argv[i], flag) == argv[i])
    def parse_args(self, argv):
        rest = []
        i = 0
        argv = list(argv)
        while i < len(argv):
            a = argv[i]
            has_next = i + 1 < len(argv)

            def nxt():
                return argv[i + 1]
            if a.startswith("--conf-file") and has_next:
                self.conf_fname = nxt(); i += 2
            elif a.startswith("--geom-file") and has_next:
                self.geom_fname = nxt().strip(" \t\n"); i += 2
            elif a.startswith("--out-dir") and has_next:
                self.out_dir = nxt().strip(" \t\n"); i += 2
            elif a.startswith("--grid-res") and has_next:
                self.resolution = _strtod(nxt()); i += 2
            elif a.startswith("--grid-num") and has_next:
                self.grid_num = _strtol(nxt()); i += 2
            elif a.startswith("--length") and has_next:
                self.len = _strtod(nxt()); i += 2
            elif a.startswith("--eps1") and has_next:
                self.ambient_eps = _strtod(nxt()); i += 2
            elif a.startswith("--save-span") and has_next:
                self.save_span = _strtol(nxt()); i += 2
            elif a.startswith("-v") and has_next:
                self.verbosity = _strtol(nxt()); i += 2
            elif a.startswith("--opts") and has_next:
                self.user_opts = nxt(); i += 2
            else:
                rest.append(a); i += 1
        if self.out_dir is None:
            self.out_dir = DEFAULT_OUT_DIR
        return rest

    def grid_cells(self):
        """meep::vol3d(L, L, L, a): (int)(L*a + 0.5) pixels per direction (disp.cpp:505-509)."""
        L = 2 * (self.len / 2 + self.pml_thickness)
        return int(L * self.resolution + 0.5)


def settings_from(conf_file=None, argv=()):
    """main.cpp:13-36: CLI first, then the conf file, then correct_defaults again."""
    s = ParseSettings()
    s.parse_args(argv)
    s.parse_conf_file(conf_file or s.conf_fname or "params.conf")
    s.correct_defaults()
    return s

