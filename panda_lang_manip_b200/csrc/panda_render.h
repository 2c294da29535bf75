// panda_render.h -- host/device structs of the analytic renderer (panda_render.cu).
#pragma once

namespace pg {

constexpr int RENDER_MAX_PRIMS = 16;
// body ids in the segmentation image: 0 = background
enum { RENDER_ID_PLANE = 1, RENDER_ID_TABLE = 2, RENDER_ID_OBJECT = 3 /* +0, +1 */, RENDER_ID_ROBOT = 5 /* base; +1+link for links 0..10 */ };

struct RenderPrim {         // 24 words
    int kind, id;           // 0 box, 1 z-cylinder, -1 end of list
    float c[3], X[3], Y[3], Z[3], h[3];   // centre, world axes of the local frame, half extents (cylinder: r, r, h/2)
    float pad[5];
};
struct RenderScene {        // what is in the picture besides the free bodies of the handle's Scene
    int has_plane, has_table, has_robot, nobj;
    float plane_z, table[4], table_height;
    float base_c[3], base_h[3];           // panda_link0 (fixed), relative to the robot base position
    float link_c[11][3], link_h[11][3];   // per link 0..10: box centre in the link frame, half extents (0 = no shape)
};
struct RenderCamera {
    int width, height, crop;
    float eye[3], fwd[3], right[3], up[3];   // camera position and orthonormal basis (world)
    float tan_half_fov, aspect, near, far;
    float light[3];
    unsigned char color[RENDER_ID_ROBOT + 1][4], background[4];
};

}  // namespace pg
