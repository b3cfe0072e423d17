// Au_graphene_box, leads and sheet ended inside the box -- the variant of junc.geom that survives a full production run.
// junc.geom (like the reference's template and its shipped Au_SiO2_box scene) runs the Au leads through the absorbing
// layer, Box([0, ...], [length, ...]).  A Drude metal (eps < 0) inside a UPML is unstable: from step ~6000 on the fields
// grow by x8.8 per 820 steps and the 31 159-step production run ends at |Ex| ~ 1e25 (scripts/blowup_bisect.py: the run
// without Au is stable, the run without graphene is not, leads ended 0.5 units inside the PML boundary decay).  Here the
// leads and the graphene sheet end 0.5 length units (5 cells) before the PML, the SiO2 substrate still fills the layer.
// Au_graphene_box -- re-authored for sim_juncs_b200 from the reference's legacy, unparseable
// junctions/Au_graphene_box/junc_template.geom (old `data(...) {}` dialect with $TOP/$BOT/$LEFT/$RGHT
// placeholders).  Geometry, polarisation and materials follow the template: two Au leads separated
// along y, a graphene sheet (make_2d) bridging the gap on top of a SiO2 substrate, Ex plane source
// at z = 1.  Concrete numbers chosen here (documented in DESIGN.md section 9):
//   gap 0.2 um, Au thickness 0.03 um, 0.75 um carrier, 1.5 fs width, 50 monitors along x at mid-gap.
// Graphene: the template's pole [1.0, 0.1, 2.88575e31, "lorentz"] diverges within a few steps
// (its ADE drive coefficient is about 1e29), so an intraband Drude sheet is used instead, with
// Fermi energy 0.4 eV and scattering time 50 fs.  The sheet plasma frequency obeys
// omega_p^2 d ~ e^2 E_F / (pi hbar^2 eps0) and sigma below is that value for a one pixel thick
// sheet at 181/18 px per unit and um_scale 16 (the make_2d rescale divides sigma by 1/resolution).
// NOTE the CGS reader splits lines at semicolons and treats `name = value` as an assignment even
// inside comments that were split that way, so comments here avoid both characters.
res = 181/18
meep_thick = um_to_l(0.03)
meep_gap = um_to_l(0.2)
mid = length/2
top = (length - meep_thick)/2
bot = (length + meep_thick)/2
left = (length - meep_gap)/2
rght = (length + meep_gap)/2
sheet_lo = (bot - 0.5/res)/res
sheet_hi = (bot + 0.5/res)/res

Gaussian_source("Ex", 0.75, 1.0, 1.5, 0.0, cutoff=4, Box([0,0,1], [length,length,1]))

monitors(locations = [vec(x, mid, top-0.1) for x in range(1,11,0.2)])

//Au: A.D. Rakic et al., Applied Optics 37, 5271 (1998), Drude term + first Lorentz term
Composite(eps = 1.0, susceptibilities = [[1e-10, 0.04274738474121455, 4.0314052191361974e21, "drude"],[0.3347200880680007, 0.19437961740816426, 11.362935694585572, "lorentz"]], [
    Box([1.5, 1.5,  top], [length-1.5, left, bot]),
    Box([1.5, rght, top], [length-1.5, length-1.5, bot])
])

//Graphene sheet, one pixel thick after the make_2d rescale (z' = z/res)
Composite(make_2d = 1, eps = 1.0, susceptibilities = [[1e-10, 0.0106, 2.398e18, "drude"]], [
    Box([1.5, left, sheet_lo], [length-1.5, rght, sheet_hi])
])

//SiO2 substrate (single Lorentz pole, HORIBA technical note), below the leads
Composite(eps = 1.0, susceptibilities = [[9.67865314895427, 0.08065544290795199, 1.12, "lorentz"]], [
    Box([0, 0, bot], [length, length, length])
])
