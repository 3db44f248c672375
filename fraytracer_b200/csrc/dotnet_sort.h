// dotnet_sort.h -- F# Array.sortInPlaceBy (SdfBoundary.fs:267-268) = .NET 6 Array.Sort(keys, items) for float keys:
// NaNs moved to the front, then introsort (insertion sort for <= 16 elements, median-of-three partition, heapsort
// once the depth budget 2*(log2 n + 1) is spent).  NOT stable: equal keys keep this algorithm's order, which is why
// it is restated operation by operation.  Host and device (the GPU grid build sorts one cell per thread).
// The recursion of the original (right part recursively, left part iteratively) is replaced by an explicit stack of
// pending sub-ranges: the sub-ranges are disjoint, so the order in which they are sorted does not matter.
#pragma once
#include <cmath>
#include <cstdint>

#include "ft_math.h"

namespace ftsort {

struct KeyItem {
    float key;
    int32_t item;
};

FT_HD void swap(KeyItem *a, int i, int j) {
    const KeyItem t = a[i];
    a[i] = a[j];
    a[j] = t;
}
FT_HD void swap_if_greater(KeyItem *a, int i, int j) {
    if (a[i].key > a[j].key) swap(a, i, j);
}
FT_HD void insertion(KeyItem *a, int n) {
    for (int i = 0; i < n - 1; i++) {
        const KeyItem t = a[i + 1];
        int j = i;
        while (j >= 0 && t.key < a[j].key) {
            a[j + 1] = a[j];
            j--;
        }
        a[j + 1] = t;
    }
}
FT_HD void down_heap(KeyItem *a, int i, int n) {
    const KeyItem d = a[i - 1];
    while (i <= n >> 1) {
        int child = 2 * i;
        if (child < n && a[child - 1].key < a[child].key) child++;
        if (!(d.key < a[child - 1].key)) break;
        a[i - 1] = a[child - 1];
        i = child;
    }
    a[i - 1] = d;
}
FT_HD void heap(KeyItem *a, int n) {
    for (int i = n >> 1; i >= 1; i--) down_heap(a, i, n);
    for (int i = n; i > 1; i--) {
        swap(a, 0, i - 1);
        down_heap(a, 1, i - 1);
    }
}
FT_HD int partition(KeyItem *a, int n) {
    const int hi = n - 1, mid = hi >> 1;
    swap_if_greater(a, 0, mid);
    swap_if_greater(a, 0, hi);
    swap_if_greater(a, mid, hi);
    const float pivot = a[mid].key;
    swap(a, mid, hi - 1);
    int left = 0, right = hi - 1;
    while (left < right) {
        while (pivot > a[++left].key) {}
        while (pivot < a[--right].key) {}
        if (left >= right) break;
        swap(a, left, right);
    }
    if (left != hi - 1) swap(a, left, hi - 1);
    return left;
}
FT_HD void intro(KeyItem *base, int n, int depth0) {
    // pending (offset, size, depth) ranges; at most one push per level of the depth budget (<= 2*(31+1))
    int st_off[72], st_n[72], st_d[72], sp = 0;
    st_off[sp] = 0; st_n[sp] = n; st_d[sp] = depth0; sp++;
    while (sp > 0) {
        sp--;
        KeyItem *a = base + st_off[sp];
        const int off = st_off[sp];
        int size = st_n[sp], depth = st_d[sp];
        while (size > 1) {
            if (size <= 16) {
                if (size == 2) {
                    swap_if_greater(a, 0, 1);
                } else if (size == 3) {
                    swap_if_greater(a, 0, 1);
                    swap_if_greater(a, 0, 2);
                    swap_if_greater(a, 1, 2);
                } else {
                    insertion(a, size);
                }
                break;
            }
            if (depth == 0) {
                heap(a, size);
                break;
            }
            depth--;
            const int p = partition(a, size);
            st_off[sp] = off + p + 1; st_n[sp] = size - (p + 1); st_d[sp] = depth; sp++;  // the right part, later
            size = p;
        }
    }
}
FT_HD void sort(KeyItem *a, int n) {
    if (n < 2) return;
    int left = 0;
    for (int i = 0; i < n; i++)
        if (a[i].key != a[i].key) {  // NaN
            swap(a, left, i);
            left++;
        }
    if (left == n) return;
    a += left;
    n -= left;
    int lg = 0;
    for (unsigned v = (unsigned)n; v >>= 1;) lg++;
    intro(a, n, 2 * (lg + 1));
}

}  // namespace ftsort
