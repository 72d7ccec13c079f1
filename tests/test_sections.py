"""Row f4: meshes of several sections -- element kinds / materials (`*Solid Section`).

CPU side: the reader, the host logic of `System_of_equations` over `SectionedFakeContext` (oracle-backed C-ABI with the
library's select / park semantics), and the KERNEL SOURCE on the SIMT emulation (multi-section pattern build with entry-id
offsets + one scatter pass per section) against the NumPy oracle.  The hardware parity tests are in test_gpu_sections.py.
The reference rejects such decks (`/root/reference/reader/inp_info.py:125-128`; first material only, `main.py:24`), so the
oracle is the sum of the reference's single-kind statements over the sections."""
import numpy as np
import pytest

from helpers import material_oracle_args, rel_err, sectioned_K, sectioned_direct_solution

from femcy_b200 import InpInfo, meshgen
from oracle import femcy_oracle as O

KINDS = ["plate_linear", "plate_quadratic", "bar_bimaterial", "bar_mixed"]


@pytest.mark.parametrize("kind", KINDS)
def test_reader_round_trip_of_a_multi_section_deck(kind, tmp_path):
    deck = meshgen.SectionedDeck(kind, n=4)
    path = str(tmp_path / (kind + ".inp"))
    meshgen.write_inp_sections(deck, path)
    inp = InpInfo(path)
    assert np.allclose(inp.nodes, deck.nodes)
    assert [s["etype"] for s in inp.sections] == [s["etype"] for s in deck.sections]
    for a, b in zip(inp.sections, deck.sections):
        assert np.array_equal(a["elements"], b["elements"])
        assert type(a["material"]) is type(b["material"]) and np.allclose(a["material"].C, b["material"].C)
        assert a["material_name"] == b["material_name"]
    assert inp.face_sets["Surf-load"] == deck.face_sets["loaded"]
    assert sorted(inp.dirichlet_bc_info[0]["node_set"].tolist()) == sorted(deck.node_sets["fixed"].tolist())
    body, material = inp.sectioned_body()
    assert material is None and len(body.parts) == len(deck.sections)


def test_single_section_decks_stay_single(tmp_path):
    """a deck the reference accepts yields one section and a plain Body (the reference's own objects)"""
    deck = meshgen.SyntheticDeck("C3D4", n=2)
    path = str(tmp_path / "one.inp")
    meshgen.write_inp(deck, path)
    inp = InpInfo(path)
    assert len(inp.sections) == 1 and inp.sections[0]["etype"] == "C3D4"
    body, material = inp.sectioned_body()
    from femcy_b200.body import SectionedBody
    assert not isinstance(body, SectionedBody) and material is list(inp.materials.values())[0]


def test_solid_section_without_known_material_or_set_raises(tmp_path):
    deck = meshgen.SectionedDeck("bar_bimaterial", n=2)
    path = str(tmp_path / "bad.inp")
    meshgen.write_inp_sections(deck, path)
    txt = open(path).read()
    open(path, "w").write(txt.replace("material=Material-2", "material=Nope"))
    with pytest.raises(ValueError, match="unknown material"):
        InpInfo(path)
    open(path, "w").write(txt.replace("*Solid Section, elset=Set-sec2", "*Solid Section, elset=Set-none"))
    with pytest.raises(ValueError, match="unknown element set"):
        InpInfo(path)


def test_mixed_dimension_decks_are_still_rejected(tmp_path):
    path = str(tmp_path / "mixdim.inp")
    open(path, "w").write("*Node\n1, 0., 0., 0.\n2, 1., 0., 0.\n3, 0., 1., 0.\n4, 0., 0., 1.\n*Element, type=C3D4\n1, 1, 2, 3, 4\n"
                          "*Element, type=CPS3\n2, 1, 2, 3\n*Step, name=s, nlgeom=NO\n*Static\n1., 1., 1e-5, 1.\n")
    with pytest.raises(ValueError, match="multiple element types"):
        InpInfo(path)


@pytest.mark.parametrize("kind", KINDS)
def test_host_logic_on_sections_matches_the_oracle(kind, monkeypatch):
    """System_of_equations over the oracle-backed context: section upload, Neumann vector over the sections, linear solve,
    per-section stress recovery, elastic energy."""
    import femcy_b200.stiffnessMtrx as sm
    from fake_ctx import SectionedFakeContext
    monkeypatch.setattr(sm, "Context", SectionedFakeContext)
    deck = meshgen.SectionedDeck(kind, n=4)
    s = sm.System_of_equations(deck.body(), None, False, quiet=True)
    assert s.sectioned and s.ctx.call("femcy_section_count") == len(deck.sections)
    s.solve(deck)
    dm = deck.nodes.shape[1]
    rhs = s.neumann_vector(deck.neumann_bc_info[0]["face_set"], 1.0, deck.neumann_bc_info[0]["direction"])
    load = rhs.reshape(-1, dm).sum(axis=0)
    assert np.allclose(load, [0., 1., 0.][:dm], atol=1e-13)          # traction 1 over a unit edge / face
    u_ref, _ = sectioned_direct_solution(deck, rhs)
    u = s.dof.to_numpy()
    assert rel_err(u, u_ref) < 1e-12
    s.compute_strain_stress()
    sig, mis = s.cauchy_stress.to_numpy(), s.mises_stress.to_numpy()
    assert len(sig) == len(deck.sections)
    for k, sec in enumerate(deck.sections):
        name, params, Cm = material_oracle_args(sec["material"])
        F = O.deformation_gradient(deck.nodes, sec["elements"], u, sec["etype"])
        ref = O.cauchy_stress(F, name, params, Cm, False)
        assert sig[k].shape == ref.shape and rel_err(sig[k], ref) < 1e-12
        mt = {"LinearIsotropicPlaneStrain": "planeStrain", "LinearIsotropicPlaneStress": "planeStress"}.get(name, "3d")
        assert rel_err(mis[k], O.mises(ref, mt, params[1])) < 1e-12
    assert s.get_elasEng() > 0


def test_loaded_facet_between_two_sections_is_rejected():
    deck = meshgen.SectionedDeck("bar_bimaterial", n=2)
    body = deck.body()
    from femcy_b200.neumann import neumann_vector_sections
    facs, _, _ = body.parts[0].boundary_arrays()
    mid = facs[np.all(np.abs(deck.nodes[facs, 0] - 1.0) < 1e-9, axis=1)]       # the plane where the two materials meet
    assert len(mid)
    with pytest.raises(KeyError, match="between two sections"):
        neumann_vector_sections(body, set(map(tuple, mid.tolist())), 1.0, np.array([0., 1., 0.]))
    with pytest.raises(KeyError, match="not on the boundary"):
        neumann_vector_sections(body, {(0, 1, 2 + deck.nodes.shape[0] // 2)}, 1.0, np.array([0., 1., 0.]))


def test_partition_and_sections_do_not_combine(monkeypatch):
    import femcy_b200.stiffnessMtrx as sm
    from fake_ctx import SectionedFakeContext
    monkeypatch.setattr(sm, "Context", SectionedFakeContext)
    deck = meshgen.SectionedDeck("bar_bimaterial", n=2)
    with pytest.raises(NotImplementedError):
        sm.System_of_equations(deck.body(), None, False, quiet=True, partition=object())


@pytest.mark.parametrize("kind", KINDS)
def test_emulated_kernels_assemble_a_multi_section_mesh(kind):
    """pattern.cu build_pattern_sections (k_elem_keys with entry-id offsets -> one sort -> per-section slots) and one
    k_assemble_scatter / k_assemble_scatter_warp pass per section, as kernel source on the SIMT emulation, against the
    sum of the oracle's per-section matrices, at a perturbed configuration."""
    import simt
    deck = meshgen.SectionedDeck(kind, n=4)
    nn, dm = deck.nodes.shape
    o = simt.build_pattern_sections([s["elements"] for s in deck.sections], nn)
    u = 0.01 * np.random.default_rng(3).standard_normal(nn * dm)
    val = None
    for sec, slots in zip(deck.sections, o["elem_slot_sections"]):
        assert slots.min() >= 0 and slots.max() < o["nslots"]
        dN, _ = sec["ELE"].device_tables()
        pat = simt.SectionPattern(o, slots, dm, nn)
        v, _, _ = simt.assemble_raw(simt.make_tables(sec["ELE"], sec["material"]), dN.shape, deck.nodes, sec["elements"], u, pat, variant=1)
        val = v if val is None else val + v
    K = simt.sell_to_csr(o, val, nn, nn, dm)
    Kref = sectioned_K(deck, u)
    assert K.nnz == Kref.nnz and np.array_equal(K.indices, Kref.indices) and np.array_equal(K.indptr, Kref.indptr)
    assert abs(K - Kref).max() <= 1e-12 * abs(Kref).max()


def test_headless_driver_runs_a_multi_section_deck_on_the_emulated_kernels(tmp_path, monkeypatch, capsys):
    """femcy_b200.main on a two-kind, two-material deck: reader -> SectionedBody -> solve -> per-section results, with
    every C-ABI call answered by the kernel source on the SIMT emulation"""
    import femcy_b200.stiffnessMtrx as sm
    from emu_ctx import EmuContext
    from femcy_b200 import main as driver
    monkeypatch.setattr(sm, "Context", EmuContext)
    deck = meshgen.SectionedDeck("plate_linear", n=4)
    path = str(tmp_path / "mixed.inp")
    meshgen.write_inp_sections(deck, path)
    out = driver.run(path, quiet=True, save=str(tmp_path / "out.npz"), vtk=str(tmp_path / "out.vtk"))
    rhs_sys = sm.System_of_equations(deck.body(), None, False, quiet=True)
    rhs = rhs_sys.neumann_vector(deck.neumann_bc_info[0]["face_set"], 1.0, deck.neumann_bc_info[0]["direction"])
    u_ref, _ = sectioned_direct_solution(deck, rhs)
    assert rel_err(out["dof"], u_ref) < 1e-8
    assert out["mises_0"].shape == (4, 4) and out["mises_1"].shape == (8, 1)
    assert "section 1 (CPS3, Material-2)" in capsys.readouterr().out
    saved = np.load(str(tmp_path / "out.npz"))
    assert np.array_equal(saved["dof"], out["dof"])
    # one VTK grid with both cell kinds (quads = type 9, triangles = type 5), cell field section after section
    from femcy_b200.vtk import read_vtk
    r = read_vtk(str(tmp_path / "out.vtk"))
    assert r["cell_types"].tolist() == [9] * 4 + [5] * 8
    assert [len(c) for c in r["cells"]] == [4] * 4 + [3] * 8
    assert np.array_equal(np.array(r["cells"][0]), deck.sections[0]["elements"][0]) and np.array_equal(np.array(r["cells"][-1]), deck.sections[1]["elements"][-1])
    assert np.allclose(r["cell_data"]["mises_gp_mean"], np.concatenate([out["mises_0"].mean(axis=1), out["mises_1"].mean(axis=1)]))
    assert np.allclose(r["point_data"]["U"][:, :2].reshape(-1), out["dof"])
