#include "delaunay.hpp"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <unordered_map>

namespace poppy {

void clip_points(std::vector<Point2f>& pts, int cols, int rows) {
    for (Point2f& p : pts) {
        p.x = p.x > cols ? cols - 1 : p.x;
        p.y = p.y > rows ? rows - 1 : p.y;
        p.x = p.x < 0 ? 0 : p.x;
        p.y = p.y < 0 ? 0 : p.y;
    }
}

namespace {
uint64_t point_key(Point2f p) {
    // exact float equality classes: +0 and -0 compare equal
    float x = p.x == 0.f ? 0.f : p.x, y = p.y == 0.f ? 0.f : p.y;
    uint32_t a, b;
    std::memcpy(&a, &x, 4);
    std::memcpy(&b, &y, 4);
    return ((uint64_t)a << 32) | b;
}
}  // namespace

void make_uniq(const std::vector<Point2f>& pts, std::vector<Point2f>& out, std::vector<int32_t>* first_index) {
    std::unordered_map<uint64_t, int> seen;
    seen.reserve(pts.size() * 2);
    for (size_t i = 0; i < pts.size(); ++i) {
        if (seen.emplace(point_key(pts[i]), (int)i).second) {
            out.push_back(pts[i]);
            if (first_index) first_index->push_back((int32_t)i);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
DelaunayMesh::DelaunayMesh(int width, int height) {
    const float big = 3.f * (float)(width > height ? width : height);
    top_left_ = {0.f, 0.f};
    bottom_right_ = {(float)width, (float)height};
    // slot 0 of both tables is the null element
    next_.assign(4, 0);
    org_.assign(4, 0);
    pt_.push_back({0.f, 0.f});
    first_edge_.push_back(0);
    const int a = new_vertex({big, 0.f}), b = new_vertex({0.f, big}), c = new_vertex({-big, -big});
    const int ab = new_quad(), bc = new_quad(), ca = new_quad();
    set_ends(ab, a, b);
    set_ends(bc, b, c);
    set_ends(ca, c, a);
    splice(ab, sym(ca));
    splice(bc, sym(ab));
    splice(ca, sym(bc));
    recent_ = ab;
}

int DelaunayMesh::new_quad() {
    if (free_quad_ <= 0) {
        next_.insert(next_.end(), 4, 0);
        org_.insert(org_.end(), 4, 0);
        free_quad_ = (int)(next_.size() / 4) - 1;
    }
    const int e = free_quad_ * 4;
    free_quad_ = next_[e + 1];
    next_[e] = e; next_[e + 1] = e + 3; next_[e + 2] = e + 2; next_[e + 3] = e + 1;
    org_[e] = org_[e + 1] = org_[e + 2] = org_[e + 3] = 0;
    return e;
}

void DelaunayMesh::free_quad_of(int e) {
    splice(e, oprev(e));
    const int s = sym(e);
    splice(s, oprev(s));
    const int q = e >> 2;
    next_[4 * q] = 0;
    next_[4 * q + 1] = free_quad_;
    free_quad_ = q;
}

int DelaunayMesh::new_vertex(Point2f p) {
    pt_.push_back(p);
    first_edge_.push_back(0);
    return (int)pt_.size() - 1;
}

void DelaunayMesh::splice(int a, int b) {
    int& an = next_[a];
    int& bn = next_[b];
    int& arn = next_[rot(an, 1)];
    int& brn = next_[rot(bn, 1)];
    std::swap(an, bn);
    std::swap(arn, brn);
}

void DelaunayMesh::set_ends(int e, int o, int d) {
    org_[e] = o;
    org_[sym(e)] = d;
    first_edge_[o] = e;
    first_edge_[d] = sym(e);
}

int DelaunayMesh::connect(int a, int b) {
    const int e = new_quad();
    splice(e, lnext(a));
    splice(sym(e), b);
    set_ends(e, dst(a), org(b));
    return e;
}

void DelaunayMesh::flip(int e) {
    const int s = sym(e), a = oprev(e), b = oprev(s);
    splice(e, a);
    splice(s, b);
    set_ends(e, dst(a), dst(b));
    splice(e, lnext(a));
    splice(s, lnext(b));
}

namespace {
double tri_area(Point2f a, Point2f b, Point2f c) {
    return ((double)b.x - a.x) * ((double)c.y - a.y) - ((double)b.y - a.y) * ((double)c.x - a.x);
}

int in_circle(Point2f pt, Point2f a, Point2f b, Point2f c) {
    const double eps = FLT_EPSILON * 0.125;
    double v = ((double)a.x * a.x + (double)a.y * a.y) * tri_area(b, c, pt);
    v -= ((double)b.x * b.x + (double)b.y * b.y) * tri_area(a, c, pt);
    v += ((double)c.x * c.x + (double)c.y * c.y) * tri_area(a, b, pt);
    v -= ((double)pt.x * pt.x + (double)pt.y * pt.y) * tri_area(a, b, c);
    return v > eps ? 1 : v < -eps ? -1 : 0;
}
}  // namespace

int DelaunayMesh::side_of(Point2f p, int e) const {
    const double cw = tri_area(p, pt_[dst(e)], pt_[org(e)]);
    return (cw > 0) - (cw < 0);
}

DelaunayMesh::Where DelaunayMesh::locate(Point2f p, int& out_edge, int& out_vertex) {
    out_edge = 0;
    out_vertex = 0;
    if (p.x < top_left_.x || p.y < top_left_.y || p.x >= bottom_right_.x || p.y >= bottom_right_.y) return kOutside;
    const int budget = (int)next_.size();
    int e = recent_;
    int r_cur = side_of(p, e);
    if (r_cur > 0) { e = sym(e); r_cur = -r_cur; }
    Where where = kError;
    for (int i = 0; i < budget; ++i) {
        const int on = onext(e), dp = dprev(e);
        const int r_on = side_of(p, on), r_dp = side_of(p, dp);
        if (r_dp > 0) {
            if (r_on > 0 || (r_on == 0 && r_cur == 0)) { where = kInside; break; }
            r_cur = r_on;
            e = on;
        } else if (r_on > 0) {
            if (r_dp == 0 && r_cur == 0) { where = kInside; break; }
            r_cur = r_dp;
            e = dp;
        } else if (r_cur == 0 && side_of(pt_[dst(on)], e) >= 0) {
            e = sym(e);
        } else {
            r_cur = r_on;
            e = on;
        }
    }
    recent_ = e;
    if (where == kInside) {
        const Point2f o = pt_[org(e)], d = pt_[dst(e)];
        double t1 = std::fabs(p.x - o.x);
        t1 += std::fabs(p.y - o.y);
        double t2 = std::fabs(p.x - d.x);
        t2 += std::fabs(p.y - d.y);
        double t3 = std::fabs(o.x - d.x);
        t3 += std::fabs(o.y - d.y);
        if (t1 < FLT_EPSILON) { out_vertex = org(e); return kVertex; }
        if (t2 < FLT_EPSILON) { out_vertex = dst(e); return kVertex; }
        if ((t1 < t3 || t2 < t3) && std::fabs(tri_area(p, o, d)) < FLT_EPSILON) where = kOnEdge;
    }
    if (where == kError) return kError;
    out_edge = e;
    return where;
}

int DelaunayMesh::insert(Point2f p) {
    int cur = 0, vtx = 0;
    const Where where = locate(p, cur, vtx);
    if (where == kOutside) { err_ = "point outside the subdivision rectangle (cv::Subdiv2D: StsOutOfRange)"; return -1; }
    if (where == kError) { err_ = "point location failed (cv::Subdiv2D: StsBadSize)"; return -1; }
    if (where == kVertex) return vtx;
    if (where == kOnEdge) {
        const int doomed = cur;
        recent_ = cur = oprev(cur);
        free_quad_of(doomed);
    }
    const int v = new_vertex(p);
    int base = new_quad();
    const int first = org(cur);
    set_ends(base, first, v);
    splice(base, cur);
    do {
        base = connect(cur, sym(base));
        cur = oprev(base);
    } while (dst(cur) != first);
    cur = oprev(base);
    const int budget = (int)next_.size();
    for (int i = 0; i < budget; ++i) {
        const int t = oprev(cur);
        const int t_dst = dst(t), c_org = org(cur), c_dst = dst(cur);
        if (side_of(pt_[t_dst], cur) > 0 && in_circle(pt_[c_org], pt_[t_dst], pt_[c_dst], pt_[v]) < 0) {
            flip(cur);
            cur = oprev(cur);
        } else if (c_org == first) {
            break;
        } else {
            cur = lprev(onext(cur));
        }
    }
    return v;
}

void DelaunayMesh::triangles(std::vector<int32_t>& ids) const {
    ids.clear();
    const int total = (int)next_.size();
    std::vector<char> seen(total, 0);
    for (int e = 4; e < total; e += 2) {
        if (seen[e]) continue;
        const int a = org(e);
        if (a < 4) continue;                  // null / bounding-triangle vertex: outside the rect
        const int eb = lnext(e), b = org(eb);
        if (b < 4) continue;
        const int ec = lnext(eb), c = org(ec);
        if (c < 4) continue;
        seen[e] = seen[eb] = seen[ec] = 1;
        ids.push_back(a);
        ids.push_back(b);
        ids.push_back(c);
    }
}

bool triangulate_points(std::vector<Point2f> points, int width, int height, std::vector<int32_t>& tri_idx,
                        std::string* error) {
    tri_idx.clear();
    clip_points(points, width, height);
    std::vector<Point2f> uniq;
    std::vector<int32_t> first;
    make_uniq(points, uniq, &first);
    DelaunayMesh mesh(width, height);
    std::vector<int32_t> owner(4, -1);        // mesh vertex id -> index of its first occurrence in `points`
    for (size_t i = 0; i < uniq.size(); ++i) {
        const int v = mesh.insert(uniq[i]);
        if (v < 0) {
            if (error) *error = mesh.error();
            return false;
        }
        if (v >= (int)owner.size()) owner.resize(v + 1, -1);
        if (owner[v] < 0) owner[v] = first[i];
    }
    std::vector<int32_t> ids;
    mesh.triangles(ids);
    tri_idx.resize(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) tri_idx[i] = owner[ids[i]];
    return true;
}

}  // namespace poppy
