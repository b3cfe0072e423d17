// continuous-wave drive (reference CW_source) onto a dielectric slab, for the CW path of bound_geom
mid = length/2
CW_source("Ex", 0.8, 1.0, 0.5, 6.0, slowness = 0.6, Box([0,0,1], [length,length,1]))
monitors(locations = [vec(mid, mid, 1.5), vec(mid, mid, mid), vec(mid/2, mid, 3.2)])
Composite(eps = 2.25, [
    Box([0, 0, mid], [length, length, mid+0.6])
])
