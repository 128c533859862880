"""A package named `sympy_to_c` for the UNMODIFIED reference to import (`from sympy_to_c import sympy_to_c as sp2c`,
pygent/modeling_scripts/cart_pole.py:7): put this directory's parent (pygent_b200/shims) on sys.path and the reference's
native-codegen seam binds the B200 implementation, pygent_b200.sympy_to_c."""
from . import sympy_to_c  # noqa: F401
from .sympy_to_c import convert_to_c  # noqa: F401
