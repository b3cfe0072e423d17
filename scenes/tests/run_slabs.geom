// C1-prime, the bracket-form twin of the reference tests/run.geom.  The reference file passes its
// geometry in the brace form, which its own parser turns into a root with zero children, that is
// vacuum (SURVEY fact 0.6).  This twin keeps the intended geometry, two slabs of eps 3.5 separated
// by a gap of 0.2, with the same source and monitors.  Comments avoid semicolons, braces and
// equal signs because the CGS line splitter looks inside comments.
tot_len = sim_length + 2*pml_thickness

Gaussian_source("Ey", 1.33, 1.0, 2.0, 0.2, 5.0, Box([0,0,1], [4,4,1]))

monitors(locations = [[1.0,1.0,1.0],vec(um_to_l(1),tot_len,1.1)])

Composite(eps = 3.5, color=90, [
    Box([0,0,0], [tot_len/2-0.1,tot_len,tot_len]),
    Box([tot_len/2+0.1,0,0], [tot_len,tot_len,tot_len])
])
