"""NumPy oracle of the three-point central difference differentiator.

Restates ``pararealml/operators/fdm/numerical_differentiator.py`` of the
reference: ``_derivative`` :1012-1057, ``_second_derivative`` :1059-1095,
``_add_halos_along_axis`` :1188-1242, the coordinate-system algebra of
``gradient`` :114-173, ``hessian`` :175-308, ``divergence`` :310-400,
``curl`` :402-590, ``laplacian`` :592-725, ``vector_laplacian`` :727-870,
``anti_laplacian`` :872-927 and ``_next_anti_laplacian_estimate`` :1097-1186.

Arrays are channels-last ``(*mesh, k)``.  ``pairs`` is a length-k sequence of
``(lower, upper)`` constraint pairs (or None) for ONE axis; ``dbc`` is the
``(x_dim, k)`` object array of such pairs.  Constraint objects only need
``.apply(array)`` and ``.multiply_and_add(addend, multiplier, result)``.
Test infrastructure only -- see ``oracle/__init__.py``.
"""
import numpy as np


def _cs(mesh) -> str:
    return mesh.coordinate_system_type.name


def _take(y, axis, sl):
    idx = [slice(None)] * y.ndim
    idx[axis] = sl
    return y[tuple(idx)]


def _pairs_or_none(dbc, x_dim, k):
    """reference :968-996"""
    if dbc is None:
        return np.empty((x_dim, k), dtype=object)
    if dbc.shape != (x_dim, k):
        raise ValueError(
            f"derivative boundary constraints must have shape {(x_dim, k)}, "
            f"got {dbc.shape}"
        )
    return dbc


def _check_shape(y, mesh, name="y"):
    """reference :930-947"""
    if y.shape[:-1] != mesh.vertices_shape:
        raise ValueError(
            f"{name} shape {y.shape[:-1]} must match mesh vertices shape "
            f"{mesh.vertices_shape}"
        )


def _check_vector_field(y, mesh):
    """reference :950-965"""
    _check_shape(y, mesh)
    if y.shape[-1] != mesh.dimensions:
        raise ValueError(
            f"y has {y.shape[-1]} components, mesh {mesh.dimensions} dims"
        )


def _r(mesh):
    return mesh.vertex_coordinate_grids[0][..., np.newaxis]


def _phi(mesh):
    return mesh.vertex_coordinate_grids[2][..., np.newaxis]


# ---------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------
def derivative(y, d_x, axis, pairs):
    """First derivative with ZERO ghost cells, then boundary planes are
    overwritten with the Neumann values (reference :1012-1057)."""
    if y.shape[axis] <= 2:
        raise ValueError(f"need at least 3 points along axis {axis}")
    ghost = np.zeros(_take(y, axis, slice(0, 1)).shape)
    padded = np.concatenate([ghost, y, ghost], axis=axis)
    out = (
        _take(padded, axis, slice(2, None))
        - _take(padded, axis, slice(0, -2))
    ) / (2.0 * d_x)
    for i, pair in enumerate(pairs):
        if pair is None:
            continue
        comp = out[..., i : i + 1]
        if pair[0] is not None:
            pair[0].apply(_take(comp, axis, slice(0, 1)))
        if pair[1] is not None:
            pair[1].apply(_take(comp, axis, slice(-1, None)))
    return out


def pad_with_ghosts(y, axis, d_x, pairs):
    """Ghost cells for second differences: ``y[1] - 2 d_x g_lo`` /
    ``y[-2] + 2 d_x g_hi`` where a Neumann value exists, 0 elsewhere
    (reference :1188-1242)."""
    inner_lo = _take(y, axis, slice(1, 2))
    inner_hi = _take(y, axis, slice(-2, -1))
    ghost_lo = np.zeros_like(inner_lo)
    ghost_hi = np.zeros_like(inner_hi)
    for i, pair in enumerate(pairs):
        if pair is None:
            continue
        if pair[0] is not None:
            pair[0].multiply_and_add(
                inner_lo[..., i : i + 1], -2.0 * d_x, ghost_lo[..., i : i + 1]
            )
        if pair[1] is not None:
            pair[1].multiply_and_add(
                inner_hi[..., i : i + 1], 2.0 * d_x, ghost_hi[..., i : i + 1]
            )
    return np.concatenate([ghost_lo, y, ghost_hi], axis=axis)


def second_derivative(y, d_x1, d_x2, axis1, axis2, pairs):
    """reference :1059-1095 (mixed: constrained d/d_axis1 followed by an
    unconstrained zero-ghost d/d_axis2)."""
    if axis1 != axis2:
        first = derivative(y, d_x1, axis1, pairs)
        return derivative(first, d_x2, axis2, [None] * y.shape[-1])
    if y.shape[axis1] <= 2:
        raise ValueError(f"need at least 3 points along axis {axis1}")
    padded = pad_with_ghosts(y, axis1, d_x1, pairs)
    return (
        _take(padded, axis1, slice(2, None))
        - 2.0 * _take(padded, axis1, slice(1, -1))
        + _take(padded, axis1, slice(0, -2))
    ) / (d_x1 * d_x2)


# ---------------------------------------------------------------------------
# coordinate-system algebra
# ---------------------------------------------------------------------------
def gradient(y, mesh, x_axis, dbc=None):
    """reference :114-173"""
    _check_shape(y, mesh)
    if not 0 <= x_axis < mesh.dimensions:
        raise ValueError(f"x axis {x_axis} out of range")
    dbc = _pairs_or_none(dbc, mesh.dimensions, y.shape[-1])
    d = derivative(y, mesh.d_x[x_axis], x_axis, dbc[x_axis])
    cs = _cs(mesh)
    if cs == "CARTESIAN":
        return d
    if cs == "SPHERICAL":
        if x_axis == 0:
            return d
        if x_axis == 1:
            return d / (_r(mesh) * np.sin(_phi(mesh)))
        return d / _r(mesh)
    return d / _r(mesh) if x_axis == 1 else d


def hessian(y, mesh, x_axis1, x_axis2, dbc=None):
    """reference :175-308"""
    _check_shape(y, mesh)
    if not (0 <= x_axis1 < mesh.dimensions and 0 <= x_axis2 < mesh.dimensions):
        raise ValueError(f"x axes ({x_axis1}, {x_axis2}) out of range")
    dbc = _pairs_or_none(dbc, mesh.dimensions, y.shape[-1])
    h = mesh.d_x
    d2 = second_derivative(
        y, h[x_axis1], h[x_axis2], x_axis1, x_axis2, dbc[x_axis1]
    )
    cs = _cs(mesh)
    if cs == "CARTESIAN":
        return d2

    def d1(a):
        return derivative(y, h[a], a, dbc[a])

    axes = {x_axis1, x_axis2}
    r = _r(mesh)
    if cs == "SPHERICAL":
        phi = _phi(mesh)
        if x_axis1 == 0 and x_axis2 == 0:
            return d2
        if x_axis1 == 1 and x_axis2 == 1:
            s, c = np.sin(phi), np.cos(phi)
            return (d1(0) + (d2 / s + c * d1(2)) / (r * s)) / r
        if x_axis1 == 2 and x_axis2 == 2:
            return (d2 / r + d1(0)) / r
        if axes == {0, 1}:
            return (d2 - d1(1) / r) / (r * np.sin(phi))
        if axes == {0, 2}:
            return (d2 - d1(2) / r) / r
        s, c = np.sin(phi), np.cos(phi)
        return (s * d2 - c * d1(1)) / (r * s) ** 2
    # polar / cylindrical
    if x_axis1 != 1 and x_axis2 != 1:
        return d2
    if x_axis1 == 1 and x_axis2 == 1:
        return (d2 / r + d1(0)) / r
    if axes == {0, 1}:
        return (d2 - d1(1) / r) / r
    return d2 / r


def divergence(y, mesh, dbc=None):
    """reference :310-400"""
    _check_vector_field(y, mesh)
    dbc = _pairs_or_none(dbc, mesh.dimensions, y.shape[-1])
    h = mesh.d_x

    def dd(i):
        return derivative(y[..., i : i + 1], h[i], i, dbc[i, i : i + 1])

    cs = _cs(mesh)
    if cs == "CARTESIAN":
        out = np.zeros(y.shape[:-1] + (1,))
        for i in range(y.shape[-1]):
            out += dd(i)
        return out
    r = _r(mesh)
    if cs == "SPHERICAL":
        phi = _phi(mesh)
        return dd(0) + (
            dd(2)
            + 2.0 * y[..., :1]
            + (dd(1) + np.cos(phi) * y[..., 2:]) / np.sin(phi)
        ) / r
    out = dd(0) + (y[..., :1] + dd(1)) / r
    return out if cs == "POLAR" else out + dd(2)


def curl(y, mesh, curl_ind=0, dbc=None):
    """reference :402-590"""
    _check_vector_field(y, mesh)
    if not 2 <= mesh.dimensions <= 3:
        raise ValueError("curl needs 2 or 3 spatial dimensions")
    if mesh.dimensions == 2 and curl_ind != 0:
        raise ValueError("2D curl only has component 0")
    if not 0 <= curl_ind < mesh.dimensions:
        raise ValueError(f"curl index {curl_ind} out of range")
    dbc = _pairs_or_none(dbc, mesh.dimensions, y.shape[-1])
    h = mesh.d_x

    def dd(comp, axis):
        return derivative(
            y[..., comp : comp + 1], h[axis], axis, dbc[axis, comp : comp + 1]
        )

    def yc(comp):
        return y[..., comp : comp + 1]

    cs = _cs(mesh)
    if cs == "CARTESIAN":
        if mesh.dimensions == 2 or curl_ind == 2:
            return dd(1, 0) - dd(0, 1)
        if curl_ind == 0:
            return dd(2, 1) - dd(1, 2)
        return dd(0, 2) - dd(2, 0)
    r = _r(mesh)
    if cs == "SPHERICAL":
        if curl_ind == 0:
            phi = _phi(mesh)
            return (
                dd(1, 2) + (np.cos(phi) * yc(1) - dd(2, 1)) / np.sin(phi)
            ) / r
        if curl_ind == 1:
            return dd(2, 0) + (yc(2) - dd(0, 2)) / r
        return -dd(1, 0) + (dd(0, 1) / np.sin(_phi(mesh)) - yc(1)) / r
    if cs == "POLAR" or curl_ind == 2:
        return dd(1, 0) + (yc(1) - dd(0, 1)) / r
    if curl_ind == 0:
        return dd(2, 1) / r - dd(1, 2)
    return dd(0, 2) - dd(2, 0)


def laplacian(y, mesh, dbc=None):
    """Element-wise scalar Laplacian (reference :592-725)."""
    _check_shape(y, mesh)
    dbc = _pairs_or_none(dbc, mesh.dimensions, y.shape[-1])
    h = mesh.d_x

    def d1(a):
        return derivative(y, h[a], a, dbc[a])

    def d2(a):
        return second_derivative(y, h[a], h[a], a, a, dbc[a])

    cs = _cs(mesh)
    if cs == "CARTESIAN":
        out = np.zeros_like(y)
        for a in range(y.ndim - 1):
            out += d2(a)
        return out
    r = _r(mesh)
    if cs == "SPHERICAL":
        phi = _phi(mesh)
        s, c = np.sin(phi), np.cos(phi)
        d_r, d_phi = d1(0), d1(2)
        d2_r, d2_theta, d2_phi = d2(0), d2(1), d2(2)
        return d2_r + (
            2 * d_r + (d2_phi + (c * d_phi + d2_theta / s) / s) / r
        ) / r
    d_r = d1(0)
    d2_r, d2_theta = d2(0), d2(1)
    out = d2_r + (d2_theta / r + d_r) / r
    return out if cs == "POLAR" else out + d2(2)


def vector_laplacian(y, mesh, ind, dbc=None):
    """reference :727-870"""
    _check_vector_field(y, mesh)
    if not 0 <= ind < mesh.dimensions:
        raise ValueError(f"vector Laplacian index {ind} out of range")
    dbc = _pairs_or_none(dbc, mesh.dimensions, y.shape[-1])
    h = mesh.d_x
    lap = laplacian(y[..., ind : ind + 1], mesh, dbc[:, ind : ind + 1])

    def dd(comp, axis):
        return derivative(
            y[..., comp : comp + 1], h[axis], axis, dbc[axis, comp : comp + 1]
        )

    cs = _cs(mesh)
    if cs == "CARTESIAN":
        return lap
    r = _r(mesh)
    y_r, y_t = y[..., :1], y[..., 1:2]
    if cs == "SPHERICAL":
        phi = _phi(mesh)
        s, c = np.sin(phi), np.cos(phi)
        y_p = y[..., 2:]
        if ind == 1:
            return lap - 2.0 * (
                y_r + dd(2, 2) + (c * y_p + dd(1, 1)) / s
            ) / r**2
        if ind == 2:
            return lap + 2.0 * (
                dd(0, 1) + (c * dd(2, 1) - y_t / 2.0) / s
            ) / (s * r**2)
        return lap + 2.0 * (
            dd(0, 2) - (y_p / 2.0 + c * dd(1, 1)) / s**2
        ) / r**2
    if ind == 0:
        return lap - (y_r + 2.0 * dd(1, 1)) / r**2
    if ind == 1:
        return lap - (y_t - 2.0 * dd(0, 1)) / r**2
    return lap


# ---------------------------------------------------------------------------
# Jacobi anti-Laplacian
# ---------------------------------------------------------------------------
def jacobi_step(y_hat, rhs, mesh, dbc):
    """One Jacobi sweep for ``laplacian(y) = rhs`` (reference :1097-1186)."""
    if not np.all(np.array(y_hat.shape[:-1]) > 2):
        raise ValueError("need at least 3 points along every axis")
    cs = _cs(mesh)
    h_sqr = np.square(mesh.d_x)
    r = r_sqr = phi = s = r_sqr_s_sqr = None
    if cs != "CARTESIAN":
        r = _r(mesh)
        r_sqr = r**2
        if cs == "SPHERICAL":
            phi = _phi(mesh)
            s = np.sin(phi)
            r_sqr_s_sqr = r_sqr * s**2

    acc = np.zeros_like(y_hat)
    for a, h in enumerate(mesh.d_x):
        padded = pad_with_ghosts(y_hat, a, h, dbc[a])
        prev = _take(padded, a, slice(0, -2))
        nxt = _take(padded, a, slice(2, None))
        term = (prev + nxt) / h_sqr[a]
        if cs == "CARTESIAN":
            acc += term
        elif cs == "SPHERICAL":
            if a == 0:
                acc += term + (nxt - prev) / (h * r)
            elif a == 1:
                acc += term / r_sqr_s_sqr
            else:
                acc += (term + np.cos(phi) * (nxt - prev) / (2.0 * h * s)) / r_sqr
        else:
            if a == 0:
                acc += term + (nxt - prev) / (2.0 * h * r)
            elif a == 1:
                acc += term / r_sqr
            else:
                acc += term
    acc -= rhs

    if cs == "CARTESIAN":
        return acc / (2.0 / h_sqr).sum()
    if cs == "SPHERICAL":
        return acc / (
            2.0 / h_sqr[0]
            + 2.0 / (h_sqr[1] * r_sqr_s_sqr)
            + 2.0 / (h_sqr[2] * r_sqr)
        )
    diag = 2.0 / h_sqr[0] + 2.0 / (h_sqr[1] * r_sqr)
    if cs == "CYLINDRICAL":
        diag += 2.0 / h_sqr[2]
    return acc / diag


def _apply_all(constraints, array):
    if constraints is None:
        return array
    for i, c in enumerate(constraints):
        if c is not None:
            c.apply(array[..., i : i + 1])
    return array


def anti_laplacian(
    rhs, mesh, y_constraints, dbc=None, y_init=None, tol=1e-3,
    return_sweeps=False,
):
    """Jacobi iteration until ``||y - y_old||_2 <= tol`` (reference
    :872-927); the start is ``np.random.random`` from the global stream
    unless ``y_init`` is given."""
    _check_shape(rhs, mesh, "Laplacian")
    dbc = _pairs_or_none(dbc, mesh.dimensions, rhs.shape[-1])
    if y_init is None:
        y = np.random.random(rhs.shape)
    else:
        if y_init.shape != rhs.shape:
            raise ValueError
        y = y_init
    _apply_all(y_constraints, y)
    diff = np.inf
    sweeps = 0
    while diff > tol:
        y_old = y
        y = jacobi_step(y_old, rhs, mesh, dbc)
        _apply_all(y_constraints, y)
        diff = float(np.linalg.norm(y - y_old))
        sweeps += 1
    return (y, sweeps) if return_sweeps else y
