/*
 * wrap.cpp — C wrapper around the reference's OWN ikd-Tree, compiled verbatim from
 * /root/reference/include/ikd-Tree/ikd_Tree/ikd_Tree.cpp (included by path at build time, never
 * copied into this repository).  Output goes to oracle/_ref/libikdtree_ref.so (git-ignored).
 * TEST INFRASTRUCTURE: used to pin the oracle's kNN / Add_Points restatement and as the kNN of
 * the CPU baseline ("kind": "reference").
 *
 * The reference uses it as KD_TREE<Point>(0.3, 0.6, 0.2) (src/Modules/Mapper.cpp:64-66) through
 * Build (Mapper.cpp:68-71), Add_Points (Mapper.cpp:73-76), Nearest_Search (Mapper.cpp:82-90)
 * and size (Mapper.cpp:32-34).
 */
#include <pcl/point_types.h>
#define __OBJECTS_H__          /* skip LIMO-Velo's ROS/PCL/Eigen headers (ikd_Tree.h:13-18) */
#include "ikd_Tree.cpp"        /* resolved with -I/root/reference/include/ikd-Tree/ikd_Tree   */

#include <cstdint>

typedef KD_TREE<Point> Tree;

static Tree::PointVector to_vec(const float* xyz, int64_t n) {
    Tree::PointVector v;
    v.reserve(n);
    for (int64_t i = 0; i < n; ++i) {
        Point p;
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
        p.time = 0; p.intensity = 0; p.range = 0;
        v.push_back(p);
    }
    return v;
}

extern "C" {
void* refikd_create(float delete_param, float balance_param, float box_length) {
    return new Tree(delete_param, balance_param, box_length);
}
void refikd_destroy(void* t) { delete (Tree*)t; }
void refikd_build(void* t, const float* xyz, int64_t n) { ((Tree*)t)->Build(to_vec(xyz, n)); }
int refikd_add_points(void* t, const float* xyz, int64_t n, int downsample) {
    Tree::PointVector v = to_vec(xyz, n);
    return ((Tree*)t)->Add_Points(v, downsample != 0);
}
int refikd_size(void* t) { return ((Tree*)t)->size(); }
int refikd_nearest(void* t, const float* q, int k, float* out_xyz, float* out_sqd) {
    Point p;
    p.x = q[0]; p.y = q[1]; p.z = q[2];
    p.time = 0; p.intensity = 0; p.range = 0;
    Tree::PointVector near;
    std::vector<float> d(k);
    ((Tree*)t)->Nearest_Search(p, k, near, d);
    int found = (int)near.size();
    for (int i = 0; i < found; ++i) {
        out_xyz[3 * i] = near[i].x; out_xyz[3 * i + 1] = near[i].y; out_xyz[3 * i + 2] = near[i].z;
        out_sqd[i] = d[i];
    }
    return found;
}
int64_t refikd_flatten(void* t, float* out_xyz, int64_t cap) {
    Tree* tr = (Tree*)t;
    Tree::PointVector st;
    tr->flatten(tr->Root_Node, st, NOT_RECORD);
    int64_t n = (int64_t)st.size();
    for (int64_t i = 0; i < n && i < cap; ++i) {
        out_xyz[3 * i] = st[i].x; out_xyz[3 * i + 1] = st[i].y; out_xyz[3 * i + 2] = st[i].z;
    }
    return n;
}
}
