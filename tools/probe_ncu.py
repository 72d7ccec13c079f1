"""Launch each hot kernel a few times at full size (for `ncu -k regex:...` captures)."""
import sys

sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
deck = meshgen.SyntheticDeck("C3D4", n=n, jitter=0.1)
s = System_of_equations(Body(deck.nodes, deck.eSets["C3D4"], deck.ELE), deck.materials["Elastic"], False, quiet=True)
nb = deck.neumann_bc_info[0]
s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
for variant in (1, 2, 1, 2):
    s.assembly_variant = variant
    s.assemble_stiffnessMtrx()
for bc in deck.dirichlet_bc_info:
    s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
s.solve_by_CG(eps=1e-30, max_iter=4, check_every=4, fixed_iters=True)
s.ctx.sync()
print("done", s.ctx.launches())
