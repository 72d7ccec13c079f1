"""Element plugins (same class names as /root/reference/element_zoo/__init__.py:3-8)."""
from .element_base import ElementBase
from .elements import (ELEMENT_TYPES, Element_linear_quadrilateral, Element_linear_tetrahedral,
                       Element_linear_triangular, Element_quadratic_quadrilateral,
                       Element_quadratic_tetrahedral, Element_quadratic_triangular)

__all__ = ["ElementBase", "ELEMENT_TYPES", "Element_linear_quadrilateral", "Element_linear_tetrahedral",
           "Element_linear_triangular", "Element_quadratic_quadrilateral", "Element_quadratic_tetrahedral",
           "Element_quadratic_triangular"]
