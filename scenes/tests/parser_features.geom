// exercises the scene-language features the shipped junction files do not reach
half = length/2
third = length/3
quarter = half/2
deep = 2*9/4*3
chain = 10-4-3
pw = 2^3^2
tern_a = (half > third) ? 11 : 22
tern_b = (half <= third) ? 11 : 22
flag = (deep == 1.5) ? 1 : 0
label = "a" + "b"
grid = [[x*y for x in range(1,4)] for y in range(3)]
pick = grid[2]
last = pick[-1]
steps = range(0.5, 2.75, 0.5)
span = linspace(third, half, 4)
flat = flatten([[1,2],[3,[4,5]],6])
bag = {name = "kettle", sizes = [3, 5, 8]}
from_bag = bag.sizes[1] + bag.sizes[0]
trig = sin(pi/6) + cos(pi/3)*sqrt(4) - exp(0)
undefined_thing = not_a_name + 2
sci = 1.5e-3 + 2e+2
neg = -half + 1

CW_source("Hy", 0.8, 0.5, 1.5, end_time=9.25, slowness=7, Box([0,0,2], [length,length,2]))
Gaussian_source("Ex", 0.7, 1.0, 1.25, 0.125, cutoff = 3, start_time = 0.5, Box([0,0,1], [length,length,1]))  // trailing remark
monitors(locations = [vec(half, half, z) for z in linspace(1.5, 3.5, 5)])
snapshot("/tmp/pf.pgm", [1.5, 1.5, 1.5], look=[-1,-1,-2], scale=2, up=[0,1,0], resolution=64)

/* a remark that
   spans lines */
Composite(eps = 2.25, color = 7, [
    Sphere([half, half, half], quarter),
    Complement([ Box([0,0,0], [third, third, third]) ]),
    Cylinder([half, half, 1], 1.5, 0.75, 0.25),
    Plane([0, 1, 1], half),
])
Composite(eps = 4, combine_type = "intersect", [
    Union([ Box([1,1,1], [half, half, half]), Sphere([third, third, third], 1) ]),
    Plane([0,0,1], [1,0,1], [0,1,1.5]),
    Difference([ Box([0,0,0], [length, length, length]), Sphere([half, half, half], 0.5), Cylinder([1,1,1], 1, 0.5) ])
])
Composite(eps = 1.5, susceptibilities = [[1.1, 0.2, 0.7, "lorentz"], [0.9, 0.1, 2.0, "drude"]], [
    Rotate(pi/5, [0,0,1], [ Box([half-1, half-1, 2], [half+1, half+1, 3]), Sphere([half, half, 2.5], 1.2) ]),
    Box([0, 0, 4], [length, length, 4.5]),
    Rotate(0.3, [1,1,0], [ Cylinder([half, half, 3], 1, 0.6, 0.9) ])
])
Composite(eps = 3, [ Intersect([ Box([0,0,0],[1,1,1]) ]) ]); Composite(eps = 5, [ Box([2,2,2],[3,3,3]) ])
