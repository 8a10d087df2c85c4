#!/bin/bash
# Probe the GPU box for any OpenGL 4.5 route the reference could run on ("probe the box at run time"):
# libGL / libEGL / OSMesa (llvmpipe), an X display, the NVIDIA EGL vendor library. Output is committed under profiles/.
echo "== date: $(date -u +%Y-%m-%dT%H:%M:%SZ)   host: $(uname -srm)"
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,driver_version --format=csv,noheader 2>&1 | head -2
echo "== ldconfig: GL / EGL / OSMesa / GLX / glfw / X11"
ldconfig -p 2>/dev/null | grep -i -E "libGL\.|libGLX|libEGL|libOSMesa|libglfw|libX11\.|libGLdispatch|libOpenGL|libnvidia-egl|libnvidia-gl|libgbm|swrast|llvmpipe" || echo "(none)"
echo "== files: dri drivers, EGL vendor json, OSMesa"
ls /usr/lib/x86_64-linux-gnu/dri 2>&1 | head -5
ls /usr/share/glvnd/egl_vendor.d /etc/glvnd/egl_vendor.d 2>&1 | head -6
find / -xdev \( -name "libOSMesa*" -o -name "libEGL_nvidia*" -o -name "libGLX_nvidia*" -o -name "swrast_dri.so" -o -name "libEGL_mesa*" \) 2>/dev/null | head -10
echo "== DISPLAY='${DISPLAY}'  WAYLAND_DISPLAY='${WAYLAND_DISPLAY}'  /dev/dri:"; ls /dev/dri 2>&1 | head -3
echo "== python ctypes.util.find_library"
python - <<'PY'
import ctypes.util
for n in ("GL", "EGL", "OSMesa", "glfw", "X11", "OpenGL"):
    print(f"  {n}: {ctypes.util.find_library(n)}")
try:
    import OpenGL  # noqa
    print("  PyOpenGL: importable")
except Exception as e:
    print("  PyOpenGL:", type(e).__name__, e)
PY
echo "== cmake / build prerequisites of the reference (CMakeLists.txt: OpenGL, GLEW, glfw3, TBB)"
for h in GL/gl.h GL/glew.h GLFW/glfw3.h EGL/egl.h GL/osmesa.h tbb/tbb.h; do
  f=$(find /usr/include /usr/local/include -path "*$h" 2>/dev/null | head -1); echo "  $h: ${f:-missing}"
done
echo "== verdict"
if ldconfig -p 2>/dev/null | grep -q -E "libOSMesa|libEGL\.so"; then echo "a GL entry point exists: try the reference under llvmpipe/EGL"; else echo "no libGL / libEGL / libOSMesa on this box: the reference (GLFW window + OpenGL 4.5 compute shaders, src/main.cpp:524-558) cannot create a context here; the reference arm of bench.py stays the CPU port of its shaders (kind: port), pinned bit for bit to the shaders compiled as C++ (oracle/_ref/libglsl_ref.so)"; fi
