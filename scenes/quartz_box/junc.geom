// quartz_box -- authored for sim_juncs_b200.  The reference ships junctions/quartz_box/params.conf
// only (no scene file and no quartz material anywhere, SURVEY fact 0.5).  Geometry follows the
// Au_SiO2_box junction with the substrate and gap material replaced by crystalline quartz.
// Quartz, ordinary ray, Sellmeier fit of G. Ghosh, Opt. Commun. 163, 95 (1999), valid 0.2 to 2 um,
//   n^2 - 1.28604141 equals 1.07044083 L^2/(L^2 - 0.0100585997) + 1.10202242 L^2/(L^2 - 100)
// written as two loss-free Lorentz poles with resonances 1/sqrt(0.0100585997) and 1/10 per um.
meep_thick = um_to_l(0.2)
meep_width = um_to_l(0.05)
mid = length/2
top = (length - meep_thick)/2
bot = (length + meep_thick)/2
left = (length - meep_width)/2
rght = (length + meep_width)/2

Gaussian_source("Ex", 0.76, 1.0, 1.2676, 0.0, cutoff=4, Box([0,0,2], [length,length,2]))

monitors(locations = [vec(x, mid, top+0.5*meep_thick) for x in linspace(mid - um_to_l(0.175), mid, 40)])

//Au leads, Rakic Drude term and first Lorentz term
Composite(eps = 1.0, susceptibilities = [[1e-10, 0.04274738474121455, 4.0314052191361974e21, "drude"],[0.3347200880680007, 0.19437961740816426, 11.362935694585572, "lorentz"]], [
    Box([0,    0, top], [left,   length, bot]),
    Box([rght, 0, top], [length, length, bot])
])

//quartz gap and substrate
Composite(eps = 1.28604141, susceptibilities = [[9.970751, 0.0, 1.07044083, "lorentz"],[0.1, 0.0, 1.10202242, "lorentz"]], [
    Box([left, 0, top], [rght,   length, bot]),
    Box([0,    0, bot], [length, length, length])
])
