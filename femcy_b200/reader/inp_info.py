"""Abaqus / CalculiX `.inp` front end.

Produces exactly the fields the reference reader produces
(`/root/reference/reader/inp_info.py:14-25`): 0-based `nodes`, `eSets`, the element plugin `ELE`,
`node_sets`, `ele_sets`, `face_sets`, `dirichlet_bc_info`, `neumann_bc_info`, `materials`,
`geometric_nonlinear`, `time_incs`.  The reference re-opens and re-scans the file once per
keyword family; here the file is read once into lines and each `read_*` method scans the cached
lines with the same keyword rules, including the behaviours decks rely on:

  * element type by substring match in a fixed order (so CPS6M parses as CPS6)   inp_info.py:67-75
  * node ids are re-sequenced in file order; sets are converted with a bare -1   inp_info.py:165-168,353-368
  * only assembly-level (`instance=`) *Nset/*Elset are honoured, `generate` ranges supported :142-161
  * `*Dsload set, P, p` is a traction of -p along the outward normal; longer lines are TRVEC
    (magnitude + direction)                                                      inp_info.py:258-271
  * `*Hyperelastic, neo hooke` data `c1, x`  ->  NeoHookean(C1=c1, D1=1/x)         inp_info.py:312-313
  * one element type / the first material only -- for every deck the reference accepts.  Row f4 lifts
    the reference's restriction (`inp_info.py:125-128` raises on several element types, `main.py:24` takes the first
    material): decks with several `*Element` types and / or several `*Solid Section` materials additionally get
    `sections` = [{"etype", "elements", "labels", "ELE", "material", "material_name"}], consumed by
    `body.SectionedBody`; `eSets`, `ELE` and `materials` keep the reference's meaning (first type / keyed by law)
"""
import sys

import numpy as np

from ..element_zoo import ELEMENT_TYPES
from ..material_zoo import (LinearIsotropic, LinearIsotropicPlaneStrain, LinearIsotropicPlaneStress,
                            NeoHookean)
from .inp_info_base import InpInfoBase

# order matters: first match wins (reference inp_info.py:67-69)
_TYPE_ORDER = ["C3D8", "C3D20", "C3D4", "C3D10", "B31", "C3D6", "CPS3", "CPE3", "CPE4", "CPS4",
               "CPE8", "CPS8", "CPS6", "CPE6"]
# (numbers per element record incl. the element id, node columns kept)
_RECORD = {"C3D8": (9, slice(1, 9)), "C3D20": (21, slice(1, 9)), "C3D4": (5, slice(1, 5)),
           "CPE4": (5, slice(1, 5)), "CPS4": (5, slice(1, 5)), "CPS8": (9, slice(1, 9)),
           "CPE8": (9, slice(1, 9)), "C3D10": (11, slice(1, 11)), "B31": (3, slice(1, 3)),
           "CPS3": (4, slice(1, 4)), "CPE3": (4, slice(1, 4)), "C3D6": (7, slice(1, 7)),
           "CPS6": (7, slice(1, 7)), "CPE6": (7, slice(1, 7))}


class FaceSet(set):
    """A loaded surface as the reference holds it -- a `set` of sorted global-node tuples (`inp_info.py:199-237`) -- that
    also keeps what the deck actually said: the (element, facet key index) pairs behind `elset, Sx`.  The device Neumann
    integration (femcy_neumann, row f1) takes those pairs directly, so no facet has to be searched for in the mesh."""
    ele = kid = None

    def with_pairs(self, ele, kid, nkeys):
        if len(ele):
            pair = np.unique(np.asarray(ele, dtype=np.int64) * nkeys + np.asarray(kid, dtype=np.int64))
            self.ele, self.kid = pair // nkeys, pair % nkeys
        else:
            self.ele, self.kid = np.zeros(0, np.int64), np.zeros(0, np.int64)
        return self


class InpInfo(InpInfoBase):
    def __init__(self, file) -> None:
        self._file = file
        with open(file, "r") as fh:
            self._lines = fh.readlines()
        # every line that holds a `*` (keywords, comments): the bulk blocks (*Node / *Element data) lie between two of
        # them and are parsed by NumPy in one go; `_light` = the deck without those blocks, which is all the other
        # `read_*` scans need (multi-million-line decks: seconds instead of minutes)
        self._star = [i for i, l in enumerate(self._lines) if "*" in l]
        self._light = None
        self.nodes, self.eSets = self.read_node_element(file)
        self.node_sets, self.ele_sets = self.read_set(file)
        self.face_sets = self.read_face_set(file)
        self.dirichlet_bc_info, self.neumann_bc_info = self.get_boundary_condition(file)
        self.materials = self.read_material(file)
        self.geometric_nonlinear = self.read_geometric_nonlinear(file)
        self.time_incs = self.read_time_inc(file)
        self.sections = self.read_sections(file)

    def _get_lines(self, file, light=False):
        if file == getattr(self, "_file", None):
            if light and self._light is not None:
                return self._light
            return self._lines
        with open(file, "r") as fh:
            return fh.readlines()

    @staticmethod
    def _numbers(block, dtype):
        """all numbers of a block of comma-separated data lines, flattened (continuation lines included)."""
        if not block:
            return np.zeros(0, dtype=dtype)
        return np.fromstring("".join(block).replace(",", " "), dtype=dtype, sep=" ")

    # ------------------------------------------------------------------------------------------
    def read_node_element(self, fileName):
        lines = self._get_lines(fileName)
        star = self._star if fileName == getattr(self, "_file", None) else [i for i, l in enumerate(lines) if "*" in l]
        nxt = {i: (star[k + 1] if k + 1 < len(star) else len(lines)) for k, i in enumerate(star)}
        bulk = []                                            # (first, last+1) line ranges of the bulk blocks

        # nodes: the data lines after the FIRST *Node keyword, up to the next line holding a `*`  (:28-43)
        ids, coords = np.zeros(0, dtype=np.int64), np.zeros((0, 0))
        for i in star:
            line = lines[i]
            if "*Node" in line or "*NODE" in line or "*node" in line:
                block = lines[i + 1:nxt[i]]
                if block:
                    width = len(block[0].split(","))
                    flat = self._numbers(block, np.float64).reshape(-1, width)
                    ids, coords = flat[:, 0].astype(np.int64), np.ascontiguousarray(flat[:, 1:])
                bulk.append((i + 1, nxt[i]))
                break

        # elements: every *Element block whose keyword line names a known type; blocks of one type concatenate (:45-75)
        tokens = {}
        for i in star:
            line = lines[i]
            if "*ELEMENT" in line or "*Element" in line or "*element" in line:
                for t in _TYPE_ORDER:
                    if ("TYPE=" in line or "type=" in line) and t in line:
                        tokens.setdefault(t, []).append(self._numbers(lines[i + 1:nxt[i]], np.int64))
                        bulk.append((i + 1, nxt[i]))
                        break
        if len(tokens) > 1:
            print("\033[31;1m there are multiple element types in the file, \033[0m")
            print("\033[40;33;1m {} \033[0m".format(list(tokens.keys())))
        eSets = {}
        self._elem_labels = {}
        for t, tok in tokens.items():
            if t not in _RECORD:
                print("\033[31;1m Error, element type {} is not found! \033[0m".format(t))
                sys.exit(1)
            width, cols = _RECORD[t]
            rec = np.concatenate(tok).reshape((-1, width))
            eSets[t] = rec[:, cols]
            self._elem_labels[t] = rec[:, 0].copy()

        if fileName == getattr(self, "_file", None) and self._light is None:
            keep, pos = [], 0
            for lo, hi in sorted(bulk):
                keep.extend(lines[pos:lo])
                pos = max(pos, hi)
            keep.extend(lines[pos:])
            self._light = keep

        nodes, eSets = self.sequence_order_of_body((ids, coords), eSets)
        first = list(eSets.keys())[0]
        self.ELE = ELEMENT_TYPES[first]()
        if len(eSets) != 1:
            # the reference raises here (inp_info.py:125-128); row f4 accepts the deck when every type is one this
            # library has kernels for and all of them live in the same dimension
            bad = [t for t in eSets if t not in ELEMENT_TYPES]
            dims = {ELEMENT_TYPES[t].dm for t in eSets if t in ELEMENT_TYPES}
            if bad or len(dims) != 1:
                raise ValueError("\033[31;1m multiple element types have not been supported now \033[0m")
        return nodes, eSets

    def _label_lookup(self):
        """element label -> (index of its type in eSets, row in that type's connectivity); -1 for unknown labels"""
        if not hasattr(self, "_lab_lut"):
            top = max(int(l.max()) for l in self._elem_labels.values() if l.size) + 1
            lut_t = np.full(top, -1, dtype=np.int64)
            lut_i = np.full(top, -1, dtype=np.int64)
            for k, t in enumerate(self.eSets):
                lab = self._elem_labels[t]
                lut_t[lab] = k
                lut_i[lab] = np.arange(lab.size)
            self._lab_lut = (lut_t, lut_i)
        return self._lab_lut

    # ------------------------------------------------------------------------------------------
    def read_set(self, fileName):
        node_sets, ele_sets = {}, {}
        target, name, generate = None, None, False
        for line in self._get_lines(fileName, light=True):
            if line[0:2] == "**":
                continue
            if line[0] == "*":
                parts = line.split(",")
                if parts[0] in ("*Nset", "*Elset") and "instance" in line:
                    target = node_sets if parts[0] == "*Nset" else ele_sets
                    name = parts[1].split("=")[1]
                    target[name] = set()
                    generate = "generate" in parts[-1]
                else:
                    target = None
                continue
            if target is not None:
                try:
                    data = list(map(int, line.split(",")))
                except ValueError:
                    data = list(map(int, line.split(",")[:-1]))
                if generate:
                    target[name] |= {*np.arange(data[0], data[1] + data[2], data[2])}
                else:
                    target[name] |= {*data}
        for sets in (node_sets, ele_sets):
            for k in sets:
                sets[k] = np.array([*sets[k]]) - 1
        return node_sets, ele_sets

    # ------------------------------------------------------------------------------------------
    def read_face_set(self, fileName):
        if not hasattr(self, "eSets"):
            self.nodes, self.eSets = self.read_node_element(fileName)
        raw = {}
        name = None
        for line in self._get_lines(fileName, light=True):
            if line[0:2] == "**":
                continue
            if line[0] == "*":
                parts = line.split("\n")[0].split(",")
                if parts[0] == "*Surface":
                    name = parts[2].split("=")[1]
                    raw[name] = []
                else:
                    name = None
                continue
            if name is not None:
                parts = line.split("\n")[0].split(",")
                raw[name].append((parts[0], parts[1]))

        _, ele_sets = self.read_set(fileName)
        face_sets = {}
        if len(self.eSets) > 1:
            # row f4: the elements of a surface are looked up by LABEL in the type they belong to, and every element
            # contributes the facet of its own kind
            lut_t, lut_i = self._label_lookup()
            types = list(self.eSets.keys())
            eles = [ELEMENT_TYPES[t]() for t in types]
            for sname, items in raw.items():
                faces = set()
                for eset, fnum in items:
                    f = int(fnum.split("S")[1]) - 1
                    labels = np.asarray(ele_sets[eset], dtype=np.int64) + 1
                    for k, t in enumerate(types):
                        rows = lut_i[labels[lut_t[labels] == k]]
                        if rows.size == 0:
                            continue
                        ce = self.eSets[t][rows]
                        for local in eles[k].inp_surface_num[f]:
                            faces.update(map(tuple, np.sort(ce[:, list(local)], axis=1).tolist()))
                face_sets[sname] = faces
            return face_sets
        conn = self.eSets[list(self.eSets.keys())[0]]
        face2node = self.ELE.inp_surface_num
        key_index = {tuple(k): i for i, k in enumerate(self.ELE.element_facets())}
        for sname, items in raw.items():
            faces = FaceSet()
            ele, kid = [], []
            for eset, fnum in items:
                f = int(fnum.split("S")[1]) - 1
                rows = np.asarray(ele_sets[eset], dtype=np.int64)
                ce = conn[rows]
                for local in face2node[f]:
                    faces.update(map(tuple, np.sort(ce[:, list(local)], axis=1).tolist()))
                    ele.append(rows)
                    kid.append(np.full(rows.size, key_index[tuple(sorted(local))], dtype=np.int64))
            face_sets[sname] = faces.with_pairs(np.concatenate(ele) if ele else [], np.concatenate(kid) if kid else [],
                                                len(key_index))
        return face_sets

    # ------------------------------------------------------------------------------------------
    def get_boundary_condition(self, fileName):
        if not hasattr(self, "node_sets"):
            self.node_sets, self.ele_sets = self.read_set(fileName)
        if not hasattr(self, "face_sets"):
            self.face_sets = self.read_face_set(fileName)
        lines = self._get_lines(fileName, light=True)

        dirichlet = []
        reading, user = False, False
        for line in lines:
            if line[0:2] == "**":
                continue
            if line[0] == "*":
                reading = line[0:9] == "*Boundary"
                user = reading and "user" in line
                continue
            if reading:
                p = line.split("\n")[0].split(",")
                disp = float(p[3]) if len(p) >= 4 else 0.
                dirichlet.append({"node_set": self.node_sets[p[0]], "dof": int(p[1]) - 1, "val": disp, "user": user})

        neumann = []
        reading = False
        for line in lines:
            if line[0:2] == "**":
                continue
            if line[0] == "*":
                reading = line[0:7] == "*Dsload"
                continue
            if reading:
                p = line.split("\n")[0].split(",")
                if len(p) <= 3:   # pressure: traction opposite to the outward normal
                    neumann.append({"face_set": self.face_sets[p[0]], "traction": -float(p[2])})
                else:             # TRVEC: magnitude + direction
                    neumann.append({"face_set": self.face_sets[p[0]], "traction": float(p[2]),
                                    "direction": np.array(list(map(float, p[3:6])))})
        return dirichlet, neumann

    # ------------------------------------------------------------------------------------------
    def read_material(self, fileName):
        materials = {}
        state, mtype = None, None
        for line in self._get_lines(fileName, light=True):
            if line[0:2] == "**":
                continue
            if line[0] == "*" and line[0:9] == "*Material":
                state = "header"
                continue
            if state == "header":
                mtype = line.split("*")[1].split("\n")[0]
                state = "data"
                continue
            if state == "data":
                if line[0] != "*":
                    materials[mtype] = list(map(float, line.split("\n")[0].split(",")))
                else:
                    state = None
        etype = list(self.eSets.keys())[0]
        if etype[0:3] in ("CPS", "CPE"):
            for key in materials:
                if key != "Elastic":
                    raise ValueError("only support linear elastic material for 2d element now.")
                cls = LinearIsotropicPlaneStress if etype[0:3] == "CPS" else LinearIsotropicPlaneStrain
                materials[key] = cls(modulus=materials["Elastic"][0], poisson_ratio=materials["Elastic"][1])
        elif etype[0:3] == "C3D":
            for key in materials:
                if key == "Elastic":
                    materials[key] = LinearIsotropic(modulus=materials["Elastic"][0], poisson_ratio=materials["Elastic"][1])
                elif "neo hooke" in key:
                    materials[key] = NeoHookean(C1=materials[key][0], D1=1. / materials[key][1])
                else:
                    raise ValueError("material type {} has not been supported now".format(key))
        return materials

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _make_material(etype, law, data):
        """material object of constitutive law `law` (the keyword line after *Material) for elements of type `etype`
        -- the reference's rules (inp_info.py:290-315): 2-D elements take plane stress / plane strain by their type"""
        if etype[0:3] in ("CPS", "CPE"):
            if law != "Elastic":
                raise ValueError("only support linear elastic material for 2d element now.")
            cls = LinearIsotropicPlaneStress if etype[0:3] == "CPS" else LinearIsotropicPlaneStrain
            return cls(modulus=data[0], poisson_ratio=data[1])
        if law == "Elastic":
            return LinearIsotropic(modulus=data[0], poisson_ratio=data[1])
        if "neo hooke" in law:
            return NeoHookean(C1=data[0], D1=1. / data[1])
        raise ValueError("material type {} has not been supported now".format(law))

    def read_sections(self, fileName):
        """Row f4 (no reference counterpart -- the reference reader never parses `*Solid Section`): the
        sections of the deck, one per (element type, material) pair in order of first appearance.

        `*Material, name=M` blocks are read BY NAME (the reference keys them by law, so two elastic materials collide),
        `*Solid Section, elset=S, material=M` assigns M to the elements of the part-level `*Elset, elset=S` (labels;
        `generate` ranges supported; an assembly-level set of that name serves when the part has none).  Elements that no
        section names take the first material, as the reference's driver does for all of them (main.py:24).  A deck with
        one element type and one material in use yields ONE section: exactly what the reference runs."""
        lines = self._get_lines(fileName, light=True)
        named, order = {}, []
        name, law, state = None, None, None
        for line in lines:
            if line[0:2] == "**":
                continue
            if line[0] == "*":
                if line[0:9] == "*Material":
                    name = [p.split("=")[1].strip() for p in line.split(",")[1:] if "name" in p.split("=")[0].lower()]
                    name = name[0] if name else "material-%d" % len(order)
                    state = "header"
                elif state == "header":
                    law = line.split("*")[1].split("\n")[0]
                    state = "data"
                else:
                    state = None
                continue
            if state == "data":
                named[name] = (law, list(map(float, line.split("\n")[0].split(","))))
                order.append(name)
                state = None
        # element sets by label: part-level sets first, assembly-level (`instance=`) ones as a fallback
        part_sets, asm_sets = {}, {}
        target, generate = None, False
        solid = []                                             # (elset name, material name)
        for line in lines:
            if line[0:2] == "**":
                continue
            if line[0] == "*":
                target = None
                parts = [p.strip() for p in line.split("\n")[0].split(",")]
                key = parts[0].lower()
                if key == "*elset":
                    nm = [p.split("=")[1] for p in parts[1:] if p.lower().startswith("elset")][0]
                    target = (asm_sets if "instance" in line else part_sets).setdefault(nm, [])
                    generate = any(p.lower() == "generate" for p in parts[1:])
                elif key == "*solid section":
                    kv = {p.split("=")[0].strip().lower(): p.split("=")[1].strip() for p in parts[1:] if "=" in p}
                    if "elset" in kv and "material" in kv:
                        solid.append((kv["elset"], kv["material"]))
                continue
            if target is not None:
                vals = [int(v) for v in line.split("\n")[0].split(",") if v.strip()]
                if generate:
                    target.extend(range(vals[0], vals[1] + 1, vals[2] if len(vals) > 2 else 1))
                else:
                    target.extend(vals)

        types = list(self.eSets.keys())
        first_name = order[0] if order else None
        lut_t, lut_i = self._label_lookup()
        mat_of = {t: np.full(self.eSets[t].shape[0], -1, dtype=np.int64) for t in types}   # index into `order`
        for eset, mname in solid:
            if mname not in named:
                raise ValueError("*Solid Section names the unknown material {}".format(mname))
            labels = part_sets.get(eset, asm_sets.get(eset))
            if labels is None:
                raise ValueError("*Solid Section names the unknown element set {}".format(eset))
            labels = np.asarray(labels, dtype=np.int64)
            labels = labels[(labels >= 0) & (labels < lut_t.size)]
            for k, t in enumerate(types):
                rows = lut_i[labels[lut_t[labels] == k]]
                mat_of[t][rows] = order.index(mname)
        sections = []
        for t in types:
            if t not in ELEMENT_TYPES:
                continue
            m = mat_of[t]
            if first_name is not None:
                m = np.where(m < 0, 0, m)
            for mi in sorted(set(m.tolist()), key=lambda v: int(np.argmax(m == v))):      # order of first appearance
                rows = np.nonzero(m == mi)[0]
                if mi < 0:
                    raise ValueError("the deck defines no material")
                law, data = named[order[mi]]
                sections.append({"etype": t, "elements": self.eSets[t][rows], "labels": self._elem_labels[t][rows],
                                 "rows": rows, "ELE": ELEMENT_TYPES[t](), "material": self._make_material(t, law, data),
                                 "material_name": order[mi]})
        return sections

    def sectioned_body(self):
        """`Body` of the deck: a plain `Body` when the deck has one section (what the reference builds, main.py:23), a
        `SectionedBody` otherwise.  Returns (body, material or None)."""
        from ..body import Body, SectionedBody
        if len(self.sections) <= 1:
            return Body(self.nodes, list(self.eSets.values())[0], self.ELE), list(self.materials.values())[0]
        return SectionedBody(self.nodes, [(s["elements"], s["ELE"], s["material"]) for s in self.sections]), None

    # ------------------------------------------------------------------------------------------
    def read_geometric_nonlinear(self, fileName) -> bool:
        for line in self._get_lines(fileName, light=True):
            if line[:5] == "*Step":
                return line.split("\n")[0].split(",")[-1].split("nlgeom=")[-1] != "NO"
        raise UnboundLocalError("no *Step keyword in the deck")  # the reference fails the same way

    def read_time_inc(self, fileName):
        reading = False
        time_incs = None
        for line in self._get_lines(fileName, light=True):
            if line[:7] == "*Static":
                reading = True
                continue
            if reading:
                if line[0:2] == "**":
                    continue
                v = list(map(float, line.split("\n")[0].split(",")))
                time_incs = {"ini_inc": v[0], "max_time": v[1], "min_inc": v[2], "max_inc": v[3]}
                break
        if time_incs["ini_inc"] > time_incs["max_inc"]:
            time_incs["ini_inc"] = time_incs["max_inc"]
        return time_incs

    @staticmethod
    def sequence_order_of_body(nodes, eSets):
        """node ids -> 0..nn-1 in file order; connectivity renumbered accordingly.  `nodes` is the reference's
        {id: coords} dict (inp_info.py:353-368) or an (ids, coords) pair of arrays.  A node id listed twice keeps its
        first position and its last coordinates, as a dict would."""
        if isinstance(nodes, dict):
            keys = np.fromiter(nodes.keys(), dtype=np.int64, count=len(nodes))
            coords = np.array(list(nodes.values()))
        else:
            keys, coords = nodes
            if len(np.unique(keys)) != len(keys):
                d = dict(zip(keys.tolist(), coords.tolist()))
                keys = np.fromiter(d.keys(), dtype=np.int64, count=len(d))
                coords = np.array(list(d.values()))
        lut = np.full(keys.max() + 1, -1, dtype=np.int64)
        lut[keys] = np.arange(len(keys))
        out = {}
        for t, conn in eSets.items():
            out[t] = lut[conn]
        return coords, out
