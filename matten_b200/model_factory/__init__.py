from .tfn_atomic_tensor import AtomicTensorModel  # noqa: F401
from .tfn_scalar_tensor import ScalarTensorModel, create_model  # noqa: F401
