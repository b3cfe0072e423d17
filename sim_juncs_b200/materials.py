"""Host-side material table construction (mirrors reference src/disp.cpp:264-283, 529-548)."""


def materials_from_regions(ambient_eps, region_eps, region_poles, max_poles=4):
    """Material id == region bit mask.  eps_inf follows cgs_material_function::in_bound with
    smooth_n = 0: sum over regions of def + (s_r - def) * in_r (the ambient value is counted once per
    region); the poles of every region containing the point are concatenated."""
    nreg = len(region_eps)
    mats = []
    for m in range(1 << nreg):
        ret = 0.0
        poles = []
        if nreg == 0:
            ret = ambient_eps
        for r in range(nreg):
            inside = (m >> r) & 1
            ret += ambient_eps + (region_eps[r] - ambient_eps) * inside / (0 + 1)
            if inside:
                poles += list(region_poles[r])
        if len(poles) > max_poles:
            raise ValueError("more than %d poles overlap at one point" % max_poles)
        mats.append((ret, poles))
    return mats
