// camera.h -- the process-global camera of the VolRen host (cppgl CameraImpl, cppgl/src/camera.{h,cpp}): only the state
// the renderer, the CLI and the Python module touch. `current_camera()` hands out ONE shared default camera, which is
// what the reference's bindings take the address of (bindings.cpp:186-194).
#pragma once

#include <memory>
#include <string>

#include "context.h"
#include "vmath.h"

namespace volren {

class CameraImpl {
public:
    explicit CameraImpl(const std::string& name) : name(name) { update(); }

    // camera.cpp:51-59 (perspective, non-skewed branch)
    void update() {
        dir = vmath::normalize(dir);
        up = vmath::normalize(up);
        view = vmath::lookAt(pos, pos + dir, up);
        view_normal = vmath::transpose(vmath::inverse(view));
        proj = vmath::perspective(vmath::radians(fov_degree), aspect_ratio(), near, far);
    }
    void from_lookat(const vmath::vec3& p, const vmath::vec3& lookat, const vmath::vec3& u = vmath::vec3(0, 1, 0)) {
        pos = p;
        dir = vmath::normalize(lookat - p);
        up = u;
        update();
    }
    static float aspect_ratio() {
        const vmath::ivec2 res = Context::initialized() ? Context::resolution() : vmath::ivec2(1280, 720);
        return float(res.x) / float(res.y);
    }

    const std::string name;
    vmath::vec3 pos = vmath::vec3(0, 0, 0), dir = vmath::vec3(1, 0, 0), up = vmath::vec3(0, 1, 0);
    float fov_degree = 70.f, near = 0.01f, far = 1000.f;
    vmath::mat4 view, view_normal, proj;
};

using Camera = std::shared_ptr<CameraImpl>;

inline Camera current_camera() {
    static Camera default_cam = std::make_shared<CameraImpl>("default");
    return default_cam;
}

}  // namespace volren
