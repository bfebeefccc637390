// Minimal dense 3-D array with the reference's indexing (x fastest, idx = i + W*(j + H*k),
// /root/reference/src/array3d.h:397-400), enough for FluidSimulation::setViscosity(Array3d<float>&).
#ifndef FLIPB200_ARRAY3D_H
#define FLIPB200_ARRAY3D_H
#include <stdexcept>
#include <vector>

template <class T>
class Array3d {
public:
    int width, height, depth;
    Array3d() : width(0), height(0), depth(0) {}
    Array3d(int w, int h, int d) : width(w), height(h), depth(d), _v((size_t)w * h * d) {}
    Array3d(int w, int h, int d, T fillv) : width(w), height(h), depth(d), _v((size_t)w * h * d, fillv) {}
    void fill(T v) { for (size_t i = 0; i < _v.size(); i++) _v[i] = v; }
    bool isIndexInRange(int i, int j, int k) const {
        return i >= 0 && j >= 0 && k >= 0 && i < width && j < height && k < depth;
    }
    T operator()(int i, int j, int k) const { return _v[_at(i, j, k)]; }
    T get(int i, int j, int k) const { return _v[_at(i, j, k)]; }
    void set(int i, int j, int k, T v) { _v[_at(i, j, k)] = v; }
    void add(int i, int j, int k, T v) { _v[_at(i, j, k)] += v; }
    T *getRawArray() { return _v.data(); }
    const T *getRawArray() const { return _v.data(); }
    int getNumElements() const { return (int)_v.size(); }

private:
    size_t _at(int i, int j, int k) const {
        if (!isIndexInRange(i, j, k)) throw std::out_of_range("Array3d: index out of range");
        return (size_t)i + (size_t)width * ((size_t)j + (size_t)height * (size_t)k);
    }
    std::vector<T> _v;
};
#endif
