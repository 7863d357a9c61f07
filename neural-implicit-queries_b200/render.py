"""Ray generator of /root/reference/src/render.py:17-50 (camera_ray, generate_camera_rays) and look_at
(:229-242): the INPUT generator of cast_rays (BASELINE configs 1 and 5).  Host NumPy float32.
`res_y` generalises the reference's square image (same per-axis formula) for 1920x1080."""
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
