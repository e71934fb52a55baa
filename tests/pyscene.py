"""Small procedural scenes in wire format (numpy), for unit-level parity tests through the raw C
ABI.  The BASELINE configs come from the C++ facade's scene recipes instead (kuafu_b200/host)."""
import numpy as np

from kuafu_b200 import wire


def look_at(eye, center, up):
    eye, center, up = (np.asarray(v, np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, f @ eye
    return m  # row-major math matrix


def perspective(fovy, aspect, near, far):
    t = np.tan(fovy / 2)
    m = np.zeros((4, 4))
    m[0, 0] = 1 / (aspect * t)
    m[1, 1] = 1 / t
    m[2, 2] = -(far + near) / (far - near)
    m[3, 2] = -1
    m[2, 3] = -(2 * far * near) / (far - near)
    return m


def col_major(m):
    return np.asarray(m, np.float64).T.reshape(16).astype(np.float32)


def camera(eye, front, up, w, h, fx=None, fy=None, cx=None, cy=None, aperture=0.0, focus=5.0):
    """CameraUBO the way reference camera.cpp:92-124 + scene.cpp:236-250 fill it."""
    eye = np.asarray(eye, np.float64)
    front = np.asarray(front, np.float64)
    fx = w * 0.5 if fx is None else fx
    fy = fx if fy is None else fy
    cx = w * 0.5 if cx is None else cx
    cy = h * 0.5 if cy is None else cy
    near, far = 0.1, 100.0
    view = look_at(eye, eye + front, up)
    proj = np.zeros((4, 4))
    proj[0, 0] = 2 * fx / w
    proj[1, 1] = -2 * fy / h
    proj[0, 2] = -2 * cx / w + 1
    proj[1, 2] = -2 * cy / h + 1
    proj[2, 2] = -far / (far - near)
    proj[3, 2] = -1
    proj[2, 3] = -far * near / (far - near)
    c = np.zeros((), wire.CAMERA)
    c["view"] = col_major(view)
    c["projection"] = col_major(proj)
    c["viewInverse"] = col_major(np.linalg.inv(view))
    c["projectionInverse"] = col_major(np.linalg.inv(proj))
    c["position"] = [*eye, aperture]
    c["front"] = [*front, focus]
    return c


def translate(v):
    m = np.eye(4)
    m[:3, 3] = v
    return m


def scale(s):
    s = np.broadcast_to(np.asarray(s, np.float64), (3,))
    return np.diag([*s, 1.0])


def rotate(angle, axis):
    axis = np.asarray(axis, np.float64)
    axis /= np.linalg.norm(axis)
    x, y, z = axis
    c, s = np.cos(angle), np.sin(angle)
    r = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s, 0],
                  [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s, 0],
                  [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c), 0],
                  [0, 0, 0, 1]])
    return r


def uv_sphere(stacks=12, slices=16, radius=1.0):
    vs, ns, uvs = [], [], []
    for i in range(stacks + 1):
        phi = np.pi * i / stacks
        for j in range(slices + 1):
            th = 2 * np.pi * j / slices
            n = np.array([np.sin(phi) * np.cos(th), np.sin(phi) * np.sin(th), np.cos(phi)])
            vs.append(n * radius)
            ns.append(n)
            uvs.append([j / slices, i / stacks])
    idx = []
    for i in range(stacks):
        for j in range(slices):
            a = i * (slices + 1) + j
            b = a + slices + 1
            if i != 0:
                idx.append([a, b, a + 1])
            if i != stacks - 1:
                idx.append([a + 1, b, b + 1])
    v = np.zeros(len(vs), wire.VERTEX)
    v["pos"], v["normal"], v["texCoord"] = vs, ns, uvs
    return v, np.asarray(idx, np.uint32).reshape(-1)


def quad(size=1.0):
    """Unit quad in the XY plane facing +Z."""
    v = np.zeros(4, wire.VERTEX)
    v["pos"] = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float32) * size
    v["normal"] = [0, 0, 1]
    v["texCoord"] = [[0, 0], [1, 0], [1, 1], [0, 1]]
    return v, np.array([0, 1, 2, 0, 2, 3], np.uint32)


def cube():
    vs, ns, uvs, idx = [], [], [], []
    for axis in range(3):
        for sign in (-1.0, 1.0):
            n = np.zeros(3)
            n[axis] = sign
            a = np.zeros(3)
            a[(axis + 1) % 3] = 1
            b = np.cross(n, a)
            base = len(vs)
            for (s, t) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
                vs.append(n + s * a + t * b)
                ns.append(n)
                uvs.append([(s + 1) / 2, (t + 1) / 2])
            idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    v = np.zeros(len(vs), wire.VERTEX)
    v["pos"], v["normal"], v["texCoord"] = vs, ns, uvs
    return v, np.asarray(idx, np.uint32)


def material(diffuse=(0.8, 0.8, 0.8), metallic=0.0, specular=0.5, roughness=0.5, ior=1.4,
             transmission=0.0, emission=(1, 1, 1), emission_strength=0.0, alpha=1.0,
             diffuse_tex=-1, metallic_tex=-1, roughness_tex=-1, transmission_tex=-1):
    m = np.zeros((), wire.MATERIAL)
    m["diffuse"] = [*diffuse, 0]
    m["emission"] = [*emission, emission_strength]
    m["alpha"], m["metallic"], m["specular"], m["roughness"] = alpha, metallic, specular, roughness
    m["ior"], m["transmission"] = ior, transmission
    m["diffuseTexIdx"], m["metallicTexIdx"] = diffuse_tex, metallic_tex
    m["roughnessTexIdx"], m["transmissionTexIdx"] = roughness_tex, transmission_tex
    return m


def instance(transform, geometry_index):
    i = np.zeros((), wire.INSTANCE)
    i["transform"] = col_major(transform)
    i["geometryIndex"] = geometry_index
    return i


def value_noise_texture(rng, size=64, channels=3):
    g = rng.integers(0, 256, (size // 8, size // 8, channels)).astype(np.float64)
    g = np.kron(g, np.ones((8, 8, 1)))
    t = np.full((size, size, 4), 255, np.uint8)
    t[..., :channels] = g[..., :channels].astype(np.uint8)
    if channels == 1:
        t[..., 1] = t[..., 0]
        t[..., 2] = t[..., 0]
    return t


def cube_env(size=16):
    faces = []
    for f in range(6):
        y, x = np.mgrid[0:size, 0:size]
        t = np.zeros((size, size, 4), np.uint8)
        t[..., 0] = 40 + 30 * f
        t[..., 1] = (x * 255 // size).astype(np.uint8)
        t[..., 2] = (y * 255 // size).astype(np.uint8)
        t[..., 3] = 255
        faces.append(t)
    return faces


class PyScene:
    def __init__(self):
        self.geoms = []      # (verts, idx, matIndex, opaque, hide)
        self.mats = []
        self.insts = []
        self.textures = []
        self.env = None
        self.dl = None
        self.pl = None
        self.al = None
        self.cams = []
        self.w = self.h = 0
        self.pc = None

    def add_geometry(self, v, idx, mat, opaque=True, hide=False):
        self.mats.append(mat)
        mi = np.full(idx.size, len(self.mats) - 1, np.uint32)  # the reference sizes it to nIndices
        self.geoms.append((v, idx, mi, opaque, hide))
        return len(self.geoms) - 1

    def upload(self, target):
        """target: kuafu_b200.rt.Context or oracle.Oracle (same call order, same buffers)."""
        is_rt = hasattr(target, "upload_geometry")
        for gi, (v, idx, mi, op, hide) in enumerate(self.geoms):
            (target.upload_geometry if is_rt else target.set_geometry)(gi, v, idx, mi, op, hide)
        mats = np.array(self.mats, wire.MATERIAL)
        (target.upload_materials if is_rt else target.set_materials)(mats)
        for ti, t in enumerate(self.textures):
            (target.upload_texture if is_rt else target.set_texture)(ti, t)
        if self.env is not None:
            (target.set_environment_cube if is_rt else target.set_env_cube)(self.env)
        target.set_lights(self.dl, self.pl, self.al)
        if is_rt:
            target.build_blas()
        target.set_instances(np.array(self.insts, wire.INSTANCE))
        if is_rt:
            target.build_tlas()

    def n_tris(self):
        return sum(self.geoms[int(i["geometryIndex"])][1].size // 3 for i in self.insts)


def small_scene(seed=0, w=96, h=64, spp=2, depth=6, lights="dir", n_spheres=6, env=False, textures=False,
                rr=False, glass=True, emissive=False, stacks=10, slices=14):
    """Floor + instanced spheres of assorted PrincipledBSDF materials (+ glass cube)."""
    rng = np.random.default_rng(seed)
    sc = PyScene()
    sc.w, sc.h = w, h
    if textures:
        sc.textures = [value_noise_texture(rng, 64, 3), value_noise_texture(rng, 64, 1),
                       value_noise_texture(rng, 32, 1)]
    qv, qi = quad()
    g_floor = sc.add_geometry(qv, qi, material(diffuse=(0.8, 0.8, 0.8), roughness=0.1,
                                               diffuse_tex=0 if textures else -1))
    sc.insts.append(instance(translate([0, 0, -1]) @ scale(12), g_floor))
    sv, si = uv_sphere(stacks, slices)
    kinds = [
        dict(diffuse=(0.7, 0.4, 0.1), metallic=1.0, specular=0.0, roughness=0.07),
        dict(diffuse=(0.2, 0.6, 0.9), metallic=0.0, specular=0.5, roughness=0.6),
        dict(diffuse=(1.0, 1.0, 1.0), metallic=0.0, specular=0.0, roughness=0.0, transmission=1.0, ior=1.45),
        dict(diffuse=(0.9, 0.9, 0.9), metallic=1.0, specular=1.0, roughness=0.0),
        dict(diffuse=(0.6, 0.9, 0.3), metallic=0.3, specular=0.4, roughness=0.3,
             roughness_tex=1 if textures else -1, transmission_tex=2 if textures else -1),
    ]
    if not glass:
        kinds.pop(2)
    g_sph = [sc.add_geometry(sv, si, material(**k)) for k in kinds]
    for k in range(n_spheres):
        pos = np.array([rng.uniform(-5, 5), rng.uniform(-4, 4), rng.uniform(-0.2, 2.0)])
        s = rng.uniform(0.5, 1.2)
        rot = rotate(rng.uniform(0, 6.28), rng.normal(size=3))
        sc.insts.append(instance(translate(pos) @ rot @ scale([s, s * rng.uniform(0.7, 1.3), s]), g_sph[k % len(g_sph)]))
    if glass:
        cv, ci = cube()
        g_cube = sc.add_geometry(cv, ci, material(diffuse=(1.0, 0.7, 0.7), metallic=0.1, specular=0.0,
                                                  roughness=0.01, ior=1.45, transmission=1.0))
        sc.insts.append(instance(translate([-1, -2, 1.5]) @ rotate(0.5, [0, 0, 1]) @ scale(0.9), g_cube))
    if emissive:
        g_em = sc.add_geometry(qv, qi, material(emission=(1.0, 0.9, 0.8), emission_strength=6.0))
        sc.insts.append(instance(translate([0, 0, 6]) @ rotate(np.pi, [1, 0, 0]) @ scale(2.0), g_em))
    if "dir" in lights:
        d = np.zeros((), wire.DIRECTIONAL_LIGHT)
        dirv = np.array([-2.0, -1.0, -1.0])
        d["direction"] = [*(dirv / np.linalg.norm(dirv)), 0.5]
        d["rgbs"] = [1.0, 0.9, 0.7, 8.0]
        sc.dl = d
    if "point" in lights:
        p = np.zeros((), wire.POINT_LIGHTS)
        p["posr"][0] = [0, 0, 4, 0.5]
        p["rgbs"][0] = [1.0, 0.5, 0.5, 100.0]
        p["posr"][1] = [-6, 0, 3, 0.0]
        p["rgbs"][1] = [0.5, 0.5, 1.0, 60.0]
        sc.pl = p
    if "active" in lights:
        a = np.zeros((), wire.ACTIVE_LIGHTS)
        view = look_at([-3.0, -3.0, 8.0], [0, 0, 0], [-1.0, 0.5, 0])
        vinv = np.linalg.inv(view)
        a["viewMat"][0] = col_major(view)
        a["projMat"][0] = col_major(perspective(np.radians(150.0), 1.0, 0.01, 1000.0))
        a["front"][0] = [-vinv[0, 2], -vinv[1, 2], -vinv[2, 2], 1]
        a["rgbs"][0] = [1, 1, 1, 1000.0]
        a["position"][0] = [vinv[0, 3], vinv[1, 3], vinv[2, 3], 0]
        a["sftp"][0] = [0.0, np.radians(150.0), 0 if textures else -1, 0]
        sc.al = a
    if env:
        sc.env = cube_env(16)
    sc.cams = [camera([-12.6, 0.0, 8.4], [0.67, 0.0, -0.5], [0, 0, 1], w, h)]
    sc.pc = wire.push_constants(clear_color=(0.64, 0.60, 0.52, 0.3), spp=spp, max_depth=depth,
                                use_env=env, rr=rr, rr_min=2)
    return sc
