from .user_api import user_dirichletBC

__all__ = ["user_dirichletBC"]
