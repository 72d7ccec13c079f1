"""Reader plugin surface.

A reader turns an input deck into the fields the solver consumes.  The reference fixes that surface
as an abstract class with eight `read_*` hooks (`/root/reference/reader/inp_info_base.py:8-40`); the
same names are required here, but they are checked when a subclass is *defined* rather than declared
one by one, and the data contract is spelled out.

Fields a reader instance must carry after construction:

    nodes                  float64 [nn, dm]      0-based, in file order
    eSets                  {element type: int [ne, n_en]}   exactly one type
    ELE                    element plugin instance (femcy_b200.element_zoo)
    node_sets, ele_sets    {name: int array}     0-based
    face_sets              {name: set of sorted global-node tuples}
    dirichlet_bc_info      [{"node_set", "dof", "val", "user"}]
    neumann_bc_info        [{"face_set", "traction"[, "direction"]}]
    materials              {keyword: material plugin instance}
    geometric_nonlinear    bool
    time_incs              {"ini_inc", "max_time", "min_inc", "max_inc"}
"""

REQUIRED_HOOKS = ("read_node_element", "read_set", "read_face_set", "get_boundary_condition",
                  "read_material", "read_geometric_nonlinear", "read_time_inc")


class InpInfoBase:
    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        missing = [h for h in REQUIRED_HOOKS if not callable(getattr(cls, h, None))]
        if missing:
            raise TypeError(f"{cls.__name__} must implement: {', '.join(missing)}")

    def __init__(self, file_name: str):
        raise TypeError("InpInfoBase is an interface; instantiate a concrete reader such as InpInfo")
