import sys
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
deck = meshgen.SyntheticDeck("C3D4", n=n, jitter=0.1)
s = System_of_equations(Body(deck.nodes, deck.eSets["C3D4"], deck.ELE), deck.materials["Elastic"], False, quiet=True)
s.assemble_stiffnessMtrx()
nb = deck.neumann_bc_info[0]
s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
for bc in deck.dirichlet_bc_info:
    s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
mi = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
s.solve_by_CG(eps=1e-8, max_iter=mi, check_every=min(100, mi))
print("iters", s.last_cg_iters, s.last_cg_residuals)
