"""Camera host logic of the hot path, device-agnostic.

Mirrors (same names, argument meaning and results; the originals hard-code ``.cuda()``):
  * ``Camera``                gaussiansplatting/scene/cameras.py:17-54
  * ``MiniCam``               gaussiansplatting/scene/cameras.py:56-67 (matrix form)
  * ``MiniCamC2W``            gs_renderer.py:853-879 (c2w form used by animation.py)
  * ``getProjectionMatrix``   gaussiansplatting/utils/graphics_utils.py:73-93
  * ``fov2focal/focal2fov``   gaussiansplatting/utils/graphics_utils.py:95-99
  * ``orbit_c2w``             threestudio/data/uncond.py:376-494 (look-at pose, +z up)
Pinned against the reference's own code by tests/golden/cameras.npz (tests/test_host_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def fov2focal(fov, pixels):
    return pixels / (2 * math.tan(fov / 2))


def focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))


def getProjectionMatrix(znear, zfar, fovX, fovY):
    """OpenGL-style frustum with P[3,2]=1 (z forward); symmetric, so the x/y offsets are exactly 0."""
    half_w = math.tan(fovX / 2) * znear
    half_h = math.tan(fovY / 2) * znear
    span = zfar - znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (half_w + half_w)
    P[1, 1] = 2.0 * znear / (half_h + half_h)
    P[2, 2] = zfar / span
    P[2, 3] = -(zfar * znear) / span
    P[3, 2] = 1.0
    return P


class Camera:
    """c2w (NeRF/OpenGL convention, 4x4) + vertical FoV -> the four tensors the rasteriser reads."""

    def __init__(self, c2w, FoVy, height, width, device="cpu"):
        FoVy = float(FoVy)
        FoVx = focal2fov(fov2focal(FoVy, height), width)
        w2c = torch.inverse(torch.as_tensor(c2w, dtype=torch.float32).cpu().clone())
        # "rectify": flip y/z camera axes and the translation sign (cameras.py:27-29)
        w2c[1:3, :3] *= -1
        w2c[:3, 3] *= -1
        self.FoVx, self.FoVy = FoVx, FoVy
        self.image_height, self.image_width = int(height), int(width)
        self.zfar, self.znear = 100.0, 0.01
        self.world_view_transform = w2c.transpose(0, 1).float().contiguous().to(device)
        self.projection_matrix = getProjectionMatrix(self.znear, self.zfar, FoVx, FoVy).transpose(0, 1).float().to(device)
        self.full_proj_transform = (
            self.world_view_transform.unsqueeze(0).bmm(self.projection_matrix.unsqueeze(0))
        ).squeeze(0).float().contiguous()
        self.camera_center = self.world_view_transform.inverse()[3, :3].float().contiguous()


class MiniCam:
    def __init__(self, width, height, fovy, fovx, znear, zfar, world_view_transform, full_proj_transform):
        self.image_width, self.image_height = width, height
        self.FoVy, self.FoVx = fovy, fovx
        self.znear, self.zfar = znear, zfar
        self.world_view_transform = world_view_transform
        self.full_proj_transform = full_proj_transform
        self.camera_center = torch.inverse(self.world_view_transform)[3][:3]


def _proj_minicam(znear, zfar, fovX, fovY):
    # gs_renderer.py:837-850 (symmetric frustum written with 1/tan)
    P = torch.zeros(4, 4)
    P[0, 0] = 1 / math.tan(fovX / 2)
    P[1, 1] = 1 / math.tan(fovY / 2)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class MiniCamC2W:
    """animation.py's camera (gs_renderer.py:853-879): numpy c2w in, camera_center = -c2w[:3,3]."""

    def __init__(self, c2w, width, height, fovy, fovx, znear, zfar, device="cpu"):
        self.image_width, self.image_height = width, height
        self.FoVy, self.FoVx = fovy, fovx
        self.znear, self.zfar = znear, zfar
        c2w = np.asarray(c2w)
        w2c = np.linalg.inv(c2w)
        w2c[1:3, :3] *= -1
        w2c[:3, 3] *= -1
        self.world_view_transform = torch.tensor(w2c).transpose(0, 1).to(device)
        self.projection_matrix = _proj_minicam(znear, zfar, fovx, fovy).transpose(0, 1).to(self.world_view_transform)
        self.full_proj_transform = self.world_view_transform @ self.projection_matrix
        self.camera_center = -torch.tensor(c2w[:3, 3]).to(device)


def orbit_c2w(elevation_deg, azimuth_deg, distance, center=(0.0, 0.0, 0.0)):
    """Look-at pose on the threestudio orbit: x back, y right, z up (uncond.py:376-494)."""
    el, az = math.radians(elevation_deg), math.radians(azimuth_deg)
    pos = torch.tensor([distance * math.cos(el) * math.cos(az), distance * math.cos(el) * math.sin(az),
                        distance * math.sin(el)], dtype=torch.float32) + torch.tensor(center, dtype=torch.float32)
    ctr = torch.tensor(center, dtype=torch.float32)
    up = torch.tensor([0.0, 0.0, 1.0])
    lookat = torch.nn.functional.normalize(ctr - pos, dim=-1)
    right = torch.nn.functional.normalize(torch.linalg.cross(lookat, up), dim=-1)
    up = torch.nn.functional.normalize(torch.linalg.cross(right, lookat), dim=-1)
    c2w = torch.eye(4)
    c2w[:3, :3] = torch.stack([right, up, -lookat], dim=-1)
    c2w[:3, 3] = pos
    return c2w


def sample_orbit_cameras(n, height, width, seed=0, elevation_range=(-30.0, 30.0), azimuth_range=(-180.0, 180.0),
                         distance_range=(1.5, 2.0), fovy_range=(40.0, 70.0), device="cpu"):
    """n cameras from the training distribution (configs/test.yaml:10,17; uncond.py:325-429):
    uniform elevation, batch-stratified azimuth, uniform distance and fovy, +z up, look-at origin."""
    g = torch.Generator().manual_seed(seed)
    el = torch.rand(n, generator=g) * (elevation_range[1] - elevation_range[0]) + elevation_range[0]
    az = (torch.rand(n, generator=g) + torch.arange(n)) / n * (azimuth_range[1] - azimuth_range[0]) + azimuth_range[0]
    dist = torch.rand(n, generator=g) * (distance_range[1] - distance_range[0]) + distance_range[0]
    fovy = torch.rand(n, generator=g) * (fovy_range[1] - fovy_range[0]) + fovy_range[0]
    # matrices are built on the host (tiny 4x4 work) and moved once: building them on the GPU costs ~10 launches per camera
    cams = [Camera(orbit_c2w(float(el[i]), float(az[i]), float(dist[i])), math.radians(float(fovy[i])), height, width, device="cpu")
            for i in range(n)]
    if str(device) != "cpu":
        for c in cams:
            c.world_view_transform = c.world_view_transform.to(device)
            c.projection_matrix = c.projection_matrix.to(device)
            c.full_proj_transform = c.full_proj_transform.to(device)
            c.camera_center = c.camera_center.to(device)
    return cams

class CameraBatch:
    """All cameras of an SDS step at once: the arithmetic of `Camera` (scene/cameras.py:17-54), batched over B poses
    (three batched tensor ops instead of ~6 small kernels per camera) -- SURVEY.md 8f-1's "camera-matrix construction".

    c2w [B,4,4], fovy [B] (radians).  Attributes: world_view_transform [B,4,4], full_proj_transform [B,4,4],
    camera_center [B,3], tanfovx / tanfovy (lists of B floats), image_height, image_width -- what rasterize_views takes."""

    def __init__(self, c2w, fovy, height, width, device="cpu", znear=0.01, zfar=100.0):
        c2w = torch.as_tensor(c2w, dtype=torch.float32).cpu().clone()
        fovy = [float(f) for f in (fovy.reshape(-1).tolist() if torch.is_tensor(fovy) else list(fovy))]
        B = c2w.shape[0]
        if len(fovy) != B:
            raise ValueError("one fovy per pose")
        w2c = torch.inverse(c2w)
        w2c[:, 1:3, :3] *= -1
        w2c[:, :3, 3] *= -1
        wvt = w2c.transpose(1, 2).contiguous()
        fovx = [focal2fov(fov2focal(fy, height), width) for fy in fovy]
        proj = torch.stack([getProjectionMatrix(znear, zfar, fx, fy).transpose(0, 1) for fx, fy in zip(fovx, fovy)])
        self.world_view_transform = wvt.to(device)
        self.full_proj_transform = torch.bmm(wvt, proj).contiguous().to(device)
        self.camera_center = torch.inverse(wvt)[:, 3, :3].contiguous().to(device)
        self.FoVx, self.FoVy = fovx, fovy
        self.tanfovx = [math.tan(f * 0.5) for f in fovx]
        self.tanfovy = [math.tan(f * 0.5) for f in fovy]
        self.image_height, self.image_width = int(height), int(width)

    def __len__(self):
        return self.world_view_transform.shape[0]
