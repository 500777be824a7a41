"""Array-shape helpers with the reference's names (harl/utils/trans_tools.py:4-27); work on torch or numpy."""


def _t2n(value):
    """torch.Tensor -> numpy (detached, host)."""
    return value.detach().cpu().numpy()


def _flatten(T, N, value):
    """[T, N, ...] -> [T*N, ...]."""
    return value.reshape(T * N, *value.shape[2:])


def _sa_cast(value):
    """[T, N, ...] -> [N*T, ...] (env-major)."""
    perm = (1, 0) + tuple(range(2, value.ndim))
    v = value.permute(*perm) if hasattr(value, "permute") else value.transpose(*perm)
    return v.reshape(-1, *value.shape[2:])


def _ma_cast(value):
    """[T, N, A, ...] -> [N*A*T, ...] (env-major, then agent, then time)."""
    perm = (1, 2, 0) + tuple(range(3, value.ndim))
    v = value.permute(*perm) if hasattr(value, "permute") else value.transpose(*perm)
    return v.reshape(-1, *value.shape[3:])
