"""`sympy_to_c.sympy_to_c` as the reference imports it: re-exports pygent_b200.sympy_to_c.convert_to_c."""
from pygent_b200.sympy_to_c import CompiledExpression, convert_to_c  # noqa: F401
