/* TEST INFRASTRUCTURE ONLY: host stand-in for the thrust calls of the reference
 * (lfm_Predictors.cu:2911-2913 sort_by_key, :2916-2921 stable_sort, :2946-2948 reduce). */
#ifndef LFM_ORACLE_THRUST_SHIM_H
#define LFM_ORACLE_THRUST_SHIM_H
#include <algorithm>
#include <functional>
#include <numeric>
#include <vector>
#include <cstddef>
namespace thrust {
template <class T> struct device_ptr {
	T* p;
	explicit device_ptr(T* p_ = nullptr) : p(p_) {}
	device_ptr operator+(std::ptrdiff_t n) const { return device_ptr(p + n); }
	std::ptrdiff_t operator-(const device_ptr& o) const { return p - o.p; }
};
template <class T> using greater = std::greater<T>;
template <class T> using plus = std::plus<T>;

/* CUB radix sort on 8-bit keys is stable; so is this */
template <class K, class V>
inline void sort_by_key(device_ptr<K> kb, device_ptr<K> ke, device_ptr<V> vb)
{
	size_t n = ke - kb;
	std::vector<size_t> idx(n);
	std::iota(idx.begin(), idx.end(), (size_t)0);
	std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return kb.p[a] < kb.p[b]; });
	std::vector<K> k2(n); std::vector<V> v2(n);
	for (size_t i = 0; i < n; i++) { k2[i] = kb.p[idx[i]]; v2[i] = vb.p[idx[i]]; }
	std::copy(k2.begin(), k2.end(), kb.p); std::copy(v2.begin(), v2.end(), vb.p);
}
template <class T, class C>
inline void stable_sort(device_ptr<T> b, device_ptr<T> e, C c) { std::stable_sort(b.p, e.p, c); }
template <class T, class A, class Op>
inline A reduce(device_ptr<T> b, device_ptr<T> e, A init, Op op)
{
	for (T* q = b.p; q != e.p; ++q) init = op(init, *q);
	return init;
}
}
#endif
