"""`sr.Mesh` of the drop-in package: a batched triangle-mesh container (reference: SoftRas/mesh.py).

Holds vertices [B,V,3], faces [B,F,3] (int) and textures (surface: [B,F,T2,3]; vertex: [B,V,3]).
`face_vertices` / `surface_normals` are cached until vertices or faces are reassigned.
OBJ loading / saving (`from_obj`, `save_obj`) are host-side I/O outside the hot path: `save_obj`
writes geometry only.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import functional as srf


class Mesh(object):
    def __init__(self, vertices, faces, textures=None, texture_res=1, texture_type="surface"):
        if isinstance(vertices, np.ndarray):
            vertices = torch.from_numpy(vertices).float().cuda()
        if isinstance(faces, np.ndarray):
            faces = torch.from_numpy(faces).int().cuda()
        if vertices.dim() == 2:
            vertices = vertices[None]
        if faces.dim() == 2:
            faces = faces[None]
        self._vertices, self._faces = vertices, faces
        self.device = vertices.device
        self.texture_type = texture_type
        self.batch_size, self.num_vertices = vertices.shape[:2]
        self.num_faces = faces.shape[1]
        self._cache = {}
        self._fill_back = False
        if textures is None:  # mesh.py:44-54: white texture
            if texture_type == "surface":
                textures = torch.ones(self.batch_size, self.num_faces, texture_res ** 2, 3,
                                      dtype=torch.float32, device=self.device)
                self.texture_res = texture_res
            elif texture_type == "vertex":
                textures = torch.ones(self.batch_size, self.num_vertices, 3, dtype=torch.float32,
                                      device=self.device)
                self.texture_res = 1
            else:
                raise ValueError("texture type not applicable")
        else:
            if isinstance(textures, np.ndarray):
                textures = torch.from_numpy(textures).float().cuda()
            if textures.dim() == 3 and texture_type == "surface":
                textures = textures[None]
            if textures.dim() == 2 and texture_type == "vertex":
                textures = textures[None]
            self.texture_res = int(np.sqrt(textures.shape[2]))  # mesh.py:63
        self._textures = textures
        self._origin = (vertices, faces, textures)

    # --- geometry ---------------------------------------------------------------------------
    @property
    def vertices(self):
        return self._vertices

    @vertices.setter
    def vertices(self, v):
        self._vertices = v
        self.num_vertices = v.shape[1]
        self._cache.clear()

    @property
    def faces(self):
        return self._faces

    @faces.setter
    def faces(self, f):
        self._faces = f
        self.num_faces = f.shape[1]
        self._cache.clear()

    @property
    def textures(self):
        return self._textures

    @textures.setter
    def textures(self, t):
        self._textures = t

    @property
    def face_vertices(self):
        if "fv" not in self._cache:
            self._cache["fv"] = srf.face_vertices(self._vertices, self._faces)
        return self._cache["fv"]

    @property
    def surface_normals(self):
        if "sn" not in self._cache:  # mesh.py:112-118
            fv = self.face_vertices
            v10 = fv[:, :, 0] - fv[:, :, 1]
            v12 = fv[:, :, 2] - fv[:, :, 1]
            self._cache["sn"] = F.normalize(torch.cross(v12, v10, dim=-1), p=2, dim=2, eps=1e-6)
        return self._cache["sn"]

    @property
    def vertex_normals(self):
        if "vn" not in self._cache:
            self._cache["vn"] = srf.vertex_normals(self._vertices, self._faces)
        return self._cache["vn"]

    @property
    def face_textures(self):
        if self.texture_type == "surface":
            return self._textures
        if self.texture_type == "vertex":
            return srf.face_vertices(self._textures, self._faces)
        raise ValueError("texture type not applicable")

    def fill_back_(self):
        if not self._fill_back:
            self.faces = torch.cat((self._faces, self._faces[:, :, [2, 1, 0]]), dim=1)
            self.textures = torch.cat((self._textures, self._textures), dim=1)
            self._fill_back = True

    def reset_(self):
        self.vertices, self.faces, self.textures = self._origin
        self._fill_back = False

    # --- host I/O (outside the hot path) ----------------------------------------------------
    def save_obj(self, filename_obj, save_texture=False, texture_res_out=16):
        if self.batch_size != 1:
            raise ValueError("Could not save when batch size >= 1")
        v = self._vertices[0].detach().cpu().numpy()
        f = self._faces[0].detach().cpu().numpy()
        with open(filename_obj, "w") as fh:
            fh.write("# umr_b200 soft_renderer.Mesh.save_obj (geometry only)\n")
            for p in v:
                fh.write("v %.8f %.8f %.8f\n" % (p[0], p[1], p[2]))
            for t in f:
                fh.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))

    @classmethod
    def from_obj(cls, filename_obj, normalization=False, load_texture=False, texture_res=1,
                 texture_type="surface"):
        verts, faces = [], []
        with open(filename_obj) as fh:
            for line in fh:
                p = line.split()
                if not p:
                    continue
                if p[0] == "v":
                    verts.append([float(x) for x in p[1:4]])
                elif p[0] == "f":
                    ids = [int(x.split("/")[0]) - 1 for x in p[1:]]
                    for k in range(1, len(ids) - 1):
                        faces.append([ids[0], ids[k], ids[k + 1]])
        v = torch.tensor(verts, dtype=torch.float32).cuda()
        f = torch.tensor(faces, dtype=torch.int32).cuda()
        if normalization:  # load_obj.py: centre and scale into [-1, 1]
            v = v - v.min(0)[0][None, :]
            v = v / torch.abs(v).max()
            v = v * 2
            v = v - v.max(0)[0][None, :] / 2
        return cls(v, f, None, texture_res, texture_type)
