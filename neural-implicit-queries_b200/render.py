"""Ray generator of /root/reference/src/render.py:17-50 (camera_ray, generate_camera_rays) and look_at
(:229-242): the INPUT generator of cast_rays (BASELINE configs 1 and 5).  Host NumPy float32.
`res_y` generalises the reference's square image (same per-axis formula) for 1920x1080.
Also the direct CALLER of cast_rays: outward_normals (:53-90, finite differences: 4 GPU point
evaluations per hit), shade_image 'normal' (:160-165), tonemap_image (:152-158) and render_image (:94-150), both the ray and the frustum
branch, without matcap shading (image assets, GUI)."""
import numpy as np

import geometry


def camera_ray(look_dir, up_dir, left_dir, fov_deg_x, fov_deg_y, theta_x, theta_y):
    f32 = np.float32
    tx = (np.asarray(theta_x, f32) * np.tan(np.deg2rad(f32(fov_deg_x)) / f32(2)).astype(f32)).astype(f32)
    ty = (np.asarray(theta_y, f32) * np.tan(np.deg2rad(f32(fov_deg_y)) / f32(2)).astype(f32)).astype(f32)
    pos = (np.asarray(look_dir, f32) + np.asarray(left_dir, f32) * tx[..., None]
           + np.asarray(up_dir, f32) * ty[..., None]).astype(f32)
    return geometry.normalize(pos)


def generate_camera_rays(eye_pos, look_dir, up_dir, res=1024, fov_deg=30., res_y=None):
    f32 = np.float32
    res_y = res if res_y is None else res_y
    eye_pos = np.asarray(eye_pos, f32)
    look_dir = np.asarray(look_dir, f32)
    up_dir = np.asarray(up_dir, f32)
    cam_ax_x = np.linspace(-1., 1., res, dtype=f32)
    cam_ax_y = np.linspace(-1., 1., res_y, dtype=f32)
    cam_x, cam_y = np.meshgrid(cam_ax_x, cam_ax_y)
    cam_x = cam_x.flatten()
    cam_y = cam_y.flatten()
    up_dir = up_dir - np.dot(look_dir, up_dir).astype(f32) * look_dir
    up_dir = geometry.normalize(up_dir)
    left_dir = np.cross(look_dir, up_dir).astype(f32)
    ray_dirs = camera_ray(look_dir, up_dir, left_dir, fov_deg, fov_deg, cam_x, cam_y)
    ray_roots = np.tile(eye_pos, (ray_dirs.shape[0], 1)).astype(f32)
    return ray_roots, ray_dirs


def look_at(eye_pos, target=None, up_dir='y'):
    f32 = np.float32
    eye_pos = np.asarray(eye_pos, f32)
    if target is None:
        target = np.array((0., 0., 0.), f32)
    if isinstance(up_dir, str):
        up_dir = np.array((0., 1., 0.), f32) if up_dir == 'y' else np.array((0., 0., 1.), f32)
    look_dir = geometry.normalize(np.asarray(target, f32) - eye_pos)
    up_dir = geometry.orthogonal_dir(up_dir, look_dir)
    left_dir = np.cross(look_dir, up_dir).astype(f32)
    return look_dir, up_dir, left_dir


def outward_normals(funcs_tuple, params_tuple, hit_pos, hit_ids, eps, method='finite_differences', ctx=None):
    """src/render.py:53-90: 'tetrahedron' central differences of the function that was hit (hit_id = i -> funcs[i-1]);
    zero where nothing was hit.  The 4 samples per point are evaluated on the GPU."""
    import mlp
    if method != 'finite_differences':
        import _niq
        raise _niq.NiqError(_niq.NIQ_EUNSUPPORTED, "outward_normals: only method='finite_differences' (no autodiff on this backend)")
    f32 = np.float32
    hit_pos = np.ascontiguousarray(hit_pos, f32)
    hit_ids = np.asarray(hit_ids)
    eps = f32(eps)
    offsets = np.array(((+eps, -eps, -eps), (-eps, -eps, +eps), (-eps, +eps, -eps), (+eps, +eps, +eps)), f32)
    x_pts = (hit_pos[:, None, :] + offsets[None, :, :]).astype(f32)
    out = np.zeros_like(hit_pos)
    for i_func, params in enumerate(params_tuple, start=1):
        samples = mlp.eval_points(params, x_pts, ctx=ctx)                                 # (N,4)
        grad = (offsets[None, :, :] * samples[:, :, None]).sum(axis=1, dtype=f32)
        with np.errstate(invalid="ignore", divide="ignore"):
            grad = geometry.normalize(grad)
        out = np.where((hit_ids == i_func)[:, None], grad, out).astype(f32)
    return out


def tonemap_image(img, gamma=2.2, white_level=.75, exposure=1.):
    """src/render.py:152-158"""
    f32 = np.float32
    img = (np.asarray(img, f32) * f32(exposure)).astype(f32)
    num = img * (f32(1.0) + (img / f32(white_level * white_level)))
    den = (f32(1.0) + img)
    return np.power((num / den).astype(f32), f32(1.0 / gamma)).astype(f32)


def shade_image(shading, ray_dirs, hit_pos, hit_normals, hit_ids, up_dir, matcaps, shading_color_tuple, shading_color_func=None):
    """src/render.py:160-224: 'normal' shading; the matcap variant needs the GUI's image assets and is not provided."""
    if shading == "normal":
        return ((np.asarray(hit_normals, np.float32) + np.float32(1.)) / np.float32(2.)).astype(np.float32)
    if shading == "matcap_color":
        import _niq
        raise _niq.NiqError(_niq.NIQ_EUNSUPPORTED, "matcap shading needs the reference GUI's image assets (outside this backend)")
    raise RuntimeError("Unrecognized shading parameter")


def render_image(funcs_tuple, params_tuple, eye_pos, look_dir, up_dir, left_dir, res, fov_deg, frustum, opts, shading="normal",
                 shading_color_tuple=((0.157, 0.613, 1.000)), matcaps=None, tonemap=False, shading_color_func=None, ctx=None):
    """src/render.py:94-150 -> (img (res,res,3), depth (res,res), counts, hit_ids, n_eval, -1)."""
    import queries
    if isinstance(funcs_tuple, list): funcs_tuple = tuple(funcs_tuple)
    if isinstance(params_tuple, list): params_tuple = tuple(params_tuple)
    if not isinstance(funcs_tuple, tuple): funcs_tuple = (funcs_tuple,)
    if not isinstance(params_tuple, tuple): params_tuple = (params_tuple,)
    if len(params_tuple) != len(funcs_tuple):
        raise ValueError("render_image tuple arguments should all be same length")
    ray_roots, ray_dirs = generate_camera_rays(eye_pos, look_dir, up_dir, res=res, fov_deg=fov_deg)
    if frustum:
        cam_params = eye_pos, look_dir, up_dir, left_dir, fov_deg, fov_deg, res, res
        t_raycast, hit_ids, counts, n_eval = queries.cast_rays_frustum(funcs_tuple, params_tuple, cam_params, opts, ctx=ctx)
        # the (res_x, res_y) images are transposed into the ray order of generate_camera_rays (src/render.py:124-126)
        t_raycast, hit_ids, counts = t_raycast.transpose().flatten(), hit_ids.transpose().flatten(), counts.transpose().flatten()
    else:
        t_raycast, hit_ids, counts, n_eval = queries.cast_rays(funcs_tuple, params_tuple, ray_roots, ray_dirs, opts, ctx=ctx)
    hit_pos = (ray_roots + t_raycast[:, None] * ray_dirs).astype(np.float32)
    hit_normals = outward_normals(funcs_tuple, params_tuple, hit_pos, hit_ids, opts['hit_eps'], ctx=ctx)
    hit_color = shade_image(shading, ray_dirs, hit_pos, hit_normals, hit_ids, up_dir, matcaps, shading_color_tuple, shading_color_func)
    img = np.where((hit_ids != 0)[:, None], hit_color, np.ones((res * res, 3), np.float32)).astype(np.float32)
    if tonemap:
        img = tonemap_image(img)
    return img.reshape(res, res, 3), t_raycast.reshape(res, res), counts.reshape(res, res), hit_ids.reshape(res, res), n_eval, -1
