"""Abstract reader surface (same members as /root/reference/reader/inp_info_base.py:8-40)."""
import abc


class InpInfoBase(abc.ABC):
    """Fields every reader must provide: nodes, eSets, ELE, node_sets, ele_sets, face_sets,
    dirichlet_bc_info, neumann_bc_info, materials, geometric_nonlinear, time_incs."""

    @abc.abstractmethod
    def __init__(self, file_name: str):
        pass

    @abc.abstractmethod
    def read_node_element(self, file_name: str):
        pass

    @abc.abstractmethod
    def read_set(self, file_name: str):
        pass

    @abc.abstractmethod
    def read_face_set(self, file_name: str):
        pass

    @abc.abstractmethod
    def get_boundary_condition(self, file_name: str):
        pass

    @abc.abstractmethod
    def read_material(self, file_name: str):
        pass

    @abc.abstractmethod
    def read_geometric_nonlinear(self, file_name: str):
        pass

    @abc.abstractmethod
    def read_time_inc(self, file_name: str):
        pass
