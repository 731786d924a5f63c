// planlab.cu -- host-only developer tool: locality statistics of candidate row orders / slot
// numberings for the level-major streaming sweep (stream.cu).  Nothing here runs on the device;
// the winner of these experiments is what pack_stream implements.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <unordered_set>

#include "hifgpu.h"

namespace hifgpu {

namespace {
unsigned lab_lanes_log2(unsigned len, unsigned U, unsigned R) {
  unsigned z = 0;
  while ((2u << z) <= R && ((len + (1u << z) - 1u) >> z) > U) ++z;
  return z;
}
}  // namespace

// opts: [0] order mode  (0 = level, length desc (round 1) ; 1 = level, class, key ; 2 = level, class, length bucket, key)
//       [1] slot mode   (0 = original ids ; 1 = level-major rank by key ; 2 = packed order)
//       [2] key mode    (0 = sweep index ; 1 = user key ; 2 = barycenter sweeps seeded by sweep index ; 3 = barycenter seeded by user key)
//       [3] barycenter iterations (forward + backward passes)
//       [4] entries per lane U (default 8)
// out:  [0] rows [1] entries [2] padded entries [3] slices [4] gather sectors (per warp instruction) [5] gather lines (128 B)
//       [6] chunk-scope sectors (distinct per 8 slices) [7] depth [8] publish sectors [9] rhs sectors
void plan_lab(const HostCsr &Tnat, bool upper, const int *user_key, const double *opts, double *out) {
  const MergeParams mp = MergeParams::from_env();
  MergeStats        ms;
  HostCsr           S = merged_sweep_form(Tnat, upper, mp, &ms);
  const unsigned    n = static_cast<unsigned>(S.nrows), m = static_cast<unsigned>(S.orig_rows);
  const int         order = static_cast<int>(opts[0]), slotmode = static_cast<int>(opts[1]),
            keymode = static_cast<int>(opts[2]), iters = static_cast<int>(opts[3]);
  const unsigned U = opts[4] > 0 ? static_cast<unsigned>(opts[4]) : 8u;
  std::vector<unsigned> lev(n, 0u);
  unsigned              depth = 0;
  for (unsigned i = 0; i < n; ++i) {
    unsigned l = 0;
    for (unsigned k = S.ptr[i]; k < S.ptr[i + 1]; ++k) l = std::max(l, lev[S.col[k]] + 1u);
    lev[i] = l;
    depth  = std::max(depth, l + 1u);
  }
  auto rowlen = [&](unsigned i) { return S.ptr[i + 1] - S.ptr[i]; };
  {  // fan-out of the solution slots: [10] most references to one row, [11] / [12] entries that reference rows
     // with >= 32 / >= 256 references, [13] rows with >= 256 references, [14] entries at <= 1 / [15] <= 4 levels distance
    std::vector<unsigned> refs(n, 0u);
    for (int c : S.col) ++refs[c];
    double mx = 0, e32 = 0, e256 = 0, r256 = 0, d1 = 0, d4 = 0;
    for (unsigned i = 0; i < n; ++i) {
      mx = std::max(mx, static_cast<double>(refs[i]));
      if (refs[i] >= 256u) r256 += 1;
    }
    for (unsigned i = 0; i < n; ++i)
      for (unsigned k = S.ptr[i]; k < S.ptr[i + 1]; ++k) {
        const unsigned c = static_cast<unsigned>(S.col[k]);
        if (refs[c] >= 32u) e32 += 1;
        if (refs[c] >= 256u) e256 += 1;
        if (lev[i] - lev[c] <= 1u) d1 += 1;
        if (lev[i] - lev[c] <= 4u) d4 += 1;
      }
    out[10] = mx, out[11] = e32, out[12] = e256, out[13] = r256, out[14] = d1, out[15] = d4;
  }
  auto orig   = [&](unsigned i) {
    const unsigned s = S.gid[i] & kCodeSlotMask;
    return s >= m ? s - m : s;
  };
  // ---- key
  std::vector<double> key(n);
  for (unsigned i = 0; i < n; ++i) {
    const unsigned o = orig(i);
    if ((keymode == 1 || keymode == 3) && user_key)
      key[i] = static_cast<double>(user_key[o]);
    else
      key[i] = static_cast<double>(upper ? m - 1u - o : o);
  }
  // rows by level
  std::vector<unsigned> lptr(depth + 1u, 0u), lrows(n);
  for (unsigned i = 0; i < n; ++i) ++lptr[lev[i] + 1u];
  for (unsigned l = 0; l < depth; ++l) lptr[l + 1] += lptr[l];
  {
    std::vector<unsigned> next(lptr.begin(), lptr.end() - 1);
    for (unsigned i = 0; i < n; ++i) lrows[next[lev[i]]++] = i;
  }
  if (keymode >= 2) {
    // layered barycenter sweeps: pos in [0,1) within each level
    std::vector<double> pos(n);
    auto rank_level = [&](unsigned l) {
      std::stable_sort(lrows.begin() + lptr[l], lrows.begin() + lptr[l + 1],
                       [&](unsigned a, unsigned b) { return key[a] < key[b]; });
      const double cnt = static_cast<double>(lptr[l + 1] - lptr[l]);
      for (unsigned q = lptr[l]; q < lptr[l + 1]; ++q) pos[lrows[q]] = (q - lptr[l] + 0.5) / cnt;
    };
    for (unsigned l = 0; l < depth; ++l) rank_level(l);
    // consumers (transpose structure)
    std::vector<unsigned> cptr(n + 1u, 0u), ccol(S.col.size());
    for (int c : S.col) ++cptr[c + 1];
    for (unsigned i = 0; i < n; ++i) cptr[i + 1] += cptr[i];
    {
      std::vector<unsigned> next(cptr.begin(), cptr.end() - 1);
      for (unsigned i = 0; i < n; ++i)
        for (unsigned k = S.ptr[i]; k < S.ptr[i + 1]; ++k) ccol[next[S.col[k]]++] = i;
    }
    for (int it = 0; it < iters; ++it) {
      // forward: a row sits at the barycenter of its producers
      for (unsigned l = 1; l < depth; ++l) {
        for (unsigned q = lptr[l]; q < lptr[l + 1]; ++q) {
          const unsigned i = lrows[q];
          double         s = 0.0;
          for (unsigned k = S.ptr[i]; k < S.ptr[i + 1]; ++k) s += pos[S.col[k]];
          key[i] = rowlen(i) ? s / rowlen(i) : pos[i];
        }
        rank_level(l);
      }
      if (it + 1 == iters) break;
      // backward: a row sits at the barycenter of its consumers
      for (unsigned l = depth - 1; l-- > 0;) {
        for (unsigned q = lptr[l]; q < lptr[l + 1]; ++q) {
          const unsigned i = lrows[q];
          double         s = 0.0;
          const unsigned c = cptr[i + 1] - cptr[i];
          for (unsigned k = cptr[i]; k < cptr[i + 1]; ++k) s += pos[ccol[k]];
          key[i] = c ? s / c : pos[i];
        }
        rank_level(l);
      }
    }
    for (unsigned i = 0; i < n; ++i) key[i] = pos[i];
  }
  // ---- class, order
  std::vector<unsigned char> cls(n);
  for (unsigned i = 0; i < n; ++i) cls[i] = static_cast<unsigned char>(lab_lanes_log2(rowlen(i), U, 32u));
  auto bucket = [&](unsigned i) -> unsigned {
    const unsigned w = (rowlen(i) + (1u << cls[i]) - 1u) >> cls[i];  // entries per lane
    return w <= 2 ? 0u : w <= 4 ? 1u : w <= 6 ? 2u : 3u;
  };
  std::vector<unsigned> ord(n);
  std::iota(ord.begin(), ord.end(), 0u);
  std::stable_sort(ord.begin(), ord.end(), [&](unsigned a, unsigned b) {
    if (lev[a] != lev[b]) return lev[a] < lev[b];
    if (order == 0) return rowlen(a) > rowlen(b);
    if (cls[a] != cls[b]) return cls[a] > cls[b];
    if (order == 2 && bucket(a) != bucket(b)) return bucket(a) > bucket(b);
    return key[a] < key[b];
  });
  // ---- slots
  std::vector<unsigned> slot(n);
  if (slotmode == 0) {
    for (unsigned i = 0; i < n; ++i) slot[i] = S.gid[i] & kCodeSlotMask;
  } else if (slotmode == 1) {
    std::vector<unsigned> o2(n);
    std::iota(o2.begin(), o2.end(), 0u);
    std::stable_sort(o2.begin(), o2.end(), [&](unsigned a, unsigned b) {
      if (lev[a] != lev[b]) return lev[a] < lev[b];
      return key[a] < key[b];
    });
    for (unsigned r = 0; r < n; ++r) slot[o2[r]] = r;
  } else {
    for (unsigned r = 0; r < n; ++r) slot[ord[r]] = r;
  }
  // ---- virtual packing + statistics
  double                       entries = 0, padded = 0, slices = 0, sectors = 0, lines = 0, chunk_sectors = 0, pub = 0, rhs_sec = 0;
  std::vector<unsigned>        ent, tmp;
  std::vector<std::vector<unsigned>> rows_cols;  // per row of the slice: slots of its entries in stored order
  std::unordered_set<unsigned> chunk_set;
  unsigned                     in_chunk = 0, chunk_level = 0xffffffffu;
  std::vector<unsigned>        inst;
  for (unsigned p = 0; p < n;) {
    const unsigned l = lev[ord[p]], z = cls[ord[p]], lpr = 1u << z, cap = 32u >> z;
    unsigned       cnt = 1;
    while (cnt < cap && p + cnt < n && lev[ord[p + cnt]] == l && cls[ord[p + cnt]] == z) ++cnt;
    unsigned width = 0;
    rows_cols.assign(cnt, {});
    for (unsigned r = 0; r < cnt; ++r) {
      const unsigned i = ord[p + r];
      width            = std::max(width, (rowlen(i) + lpr - 1u) >> z);
      std::vector<unsigned> &rc = rows_cols[r];
      for (unsigned k = S.ptr[i]; k < S.ptr[i + 1]; ++k) rc.push_back(static_cast<unsigned>(S.col[k]));
      if (slotmode == 0)
        std::stable_sort(rc.begin(), rc.end(), [&](unsigned a, unsigned b) { return lev[a] < lev[b]; });
      else
        std::stable_sort(rc.begin(), rc.end(), [&](unsigned a, unsigned b) { return slot[a] < slot[b]; });
      for (unsigned &c : rc) c = slot[c];
      entries += rc.size();
    }
    padded += static_cast<double>(width) * 32u;
    slices += 1;
    if (l != chunk_level || in_chunk == 8) {
      chunk_sectors += chunk_set.size();
      chunk_set.clear();
      in_chunk    = 0;
      chunk_level = l;
    }
    ++in_chunk;
    for (unsigned k = 0; k < width; ++k) {
      inst.clear();
      for (unsigned r = 0; r < cnt; ++r)
        for (unsigned q = 0; q < lpr; ++q) {
          const unsigned e = k * lpr + q;
          if (e < rows_cols[r].size()) inst.push_back(rows_cols[r][e]);
        }
      tmp = inst;
      for (unsigned &c : tmp) c >>= 2;
      std::sort(tmp.begin(), tmp.end());
      const unsigned ns = static_cast<unsigned>(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
      sectors += ns;
      for (unsigned q = 0; q < ns; ++q) chunk_set.insert(tmp[q]);
      for (unsigned q = 0; q < ns; ++q) tmp[q] >>= 2;
      lines += static_cast<double>(std::unique(tmp.begin(), tmp.begin() + ns) - tmp.begin());
    }
    // publish + rhs sectors of the slice
    tmp.clear();
    for (unsigned r = 0; r < cnt; ++r) tmp.push_back(slot[ord[p + r]] >> 2);
    std::sort(tmp.begin(), tmp.end());
    pub += static_cast<double>(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    tmp.clear();
    for (unsigned r = 0; r < cnt; ++r)
      if (!(S.gid[ord[p + r]] & kCodeZeroRhs)) tmp.push_back(orig(ord[p + r]) >> 2);
    std::sort(tmp.begin(), tmp.end());
    rhs_sec += static_cast<double>(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    p += cnt;
  }
  chunk_sectors += chunk_set.size();
  out[0] = n, out[1] = entries, out[2] = padded - entries, out[3] = slices, out[4] = sectors, out[5] = lines;
  out[6] = chunk_sectors, out[7] = depth, out[8] = pub, out[9] = rhs_sec;
}

}  // namespace hifgpu
