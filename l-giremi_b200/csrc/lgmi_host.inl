// lgmi_host.inl -- host-side native pieces either side of the MI step (SURVEY 8f, f1 + f2):
//
//   lgmi_cs_scan      one read's short-form cs tag -> its mismatches in contig coordinates with
//                     the splice-distance filter applied, and its introns
//                     (giremi/cs.py:8-41 cs_to_list, :573-613 get_mismatches / get_introns,
//                      giremi/mismatch.py:99-141 FILTER 1 with utils.py:4-31 interval logic)
//   lgmi_encode_unit  one (footprint, strand) unit from its flattened `mismatches[strand]`
//                     dict (read names as one newline-separated blob) -> bit-planes + flag bytes,
//                     with the dict semantics of giremi/mutual_information.py:15-16 (last allele
//                     wins), :25-32 (depth ranking, stable ties) and :33-38 (labels)
//
// Pure C++ (no CUDA): usable without a device.  Included at the end of lgmi.cu.
#include <unordered_map>
#include <string_view>

extern "C" int lgmi_cs_scan(const char* cs, uint64_t cs_len, int64_t ref_start, int min_dist_from_splice,
                            uint32_t cap_mismatch, int64_t* mm_pos, char* mm_ref, char* mm_alt, uint32_t* n_mismatch,
                            uint32_t cap_intron, int64_t* intron_lo, int64_t* intron_hi, uint32_t* n_intron) {
  if (!cs || !n_mismatch || !n_intron) return LGMI_ERR_ARG;
  std::vector<int64_t> pos_v;
  std::vector<char> ref_v, alt_v;
  std::vector<int64_t> ilo, ihi;
  int64_t pos = 0;  // relative to the alignment start (cs.py:18, 0-based)
  uint64_t k = 0;
  auto is_value = [](char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z'); };
  while (k < cs_len) {
    const char op = cs[k++];
    const uint64_t v0 = k;
    while (k < cs_len && is_value(cs[k])) ++k;
    const uint64_t vlen = k - v0;
    if (vlen == 0) return LGMI_ERR_ARG;  // the reference's zip of marks and values would go out of step
    switch (op) {
      case ':': {  // identical run: int(value)  (cs.py:21)
        int64_t n = 0;
        for (uint64_t q = v0; q < k; ++q) {
          if (cs[q] < '0' || cs[q] > '9') return LGMI_ERR_ARG;
          n = n * 10 + (cs[q] - '0');
        }
        pos += n;
        break;
      }
      case '*':  // substitution, one reference base (cs.py:22): value = ref base, read base
        if (vlen < 2) return LGMI_ERR_ARG;
        pos_v.push_back(pos);
        ref_v.push_back((char)(cs[v0] >= 'a' ? cs[v0] - 32 : cs[v0]));
        alt_v.push_back((char)(cs[v0 + 1] >= 'a' ? cs[v0 + 1] - 32 : cs[v0 + 1]));
        pos += 1;
        break;
      case '+':  // insertion: no reference advance (cs.py:23)
        break;
      case '-':  // deletion: len(value) (cs.py:24)
        pos += (int64_t)vlen;
        break;
      case '~': {  // intron: the digits inside the value (cs.py:25-27)
        int64_t n = 0;
        bool any = false;
        for (uint64_t q = v0; q < k; ++q)
          if (cs[q] >= '0' && cs[q] <= '9') {
            n = n * 10 + (cs[q] - '0');
            any = true;
          }
        if (!any) return LGMI_ERR_ARG;
        ilo.push_back(pos);
        pos += n;
        ihi.push_back(pos);
        break;
      }
      default:
        return LGMI_ERR_ARG;  // the reference raises KeyError on any other mark (cs.py:20-28)
    }
  }
  *n_intron = (uint32_t)ilo.size();
  for (uint32_t q = 0; q < ilo.size() && q < cap_intron; ++q) {
    if (intron_lo) intron_lo[q] = ilo[q] + ref_start;
    if (intron_hi) intron_hi[q] = ihi[q] + ref_start;
  }
  // FILTER 1 (mismatch.py:117-141): +-d around every intron start and end, merged (utils.py:4-16:
  // an interval joins the previous one unless it starts beyond its end), membership
  // start <= pos < end (utils.py:19-31, both searchsorted calls use side='right')
  std::vector<std::pair<int64_t, int64_t>> iv;
  if (!ilo.empty() && min_dist_from_splice > 0) {
    std::vector<int64_t> sp(ilo);
    sp.insert(sp.end(), ihi.begin(), ihi.end());
    std::sort(sp.begin(), sp.end());
    for (int64_t a : sp) {
      const int64_t lo = a - min_dist_from_splice, hi = a + min_dist_from_splice;
      if (iv.empty() || iv.back().second < lo) iv.emplace_back(lo, hi);
      else iv.back().second = std::max(iv.back().second, hi);
    }
  }
  uint32_t n = 0;
  for (size_t q = 0; q < pos_v.size(); ++q) {
    const int64_t p = pos_v[q];  // the reference filters in contig coordinates; a shift does not change membership
    bool inside = false;
    for (const auto& x : iv)
      if (x.first <= p && p < x.second) {
        inside = true;
        break;
      }
    if (inside) continue;
    if (n < cap_mismatch) {
      if (mm_pos) mm_pos[n] = p + ref_start;
      if (mm_ref) mm_ref[n] = ref_v[q];
      if (mm_alt) mm_alt[n] = alt_v[q];
    }
    ++n;
  }
  *n_mismatch = n;
  return (n > cap_mismatch || ilo.size() > cap_intron) ? LGMI_ERR_NOMEM : LGMI_OK;
}

extern "C" int lgmi_encode_unit(uint32_t n_sites, const uint8_t* site_type, const uint32_t* n_depth_entries,
                                const uint32_t* depth_allele, const int64_t* depth_value, const uint32_t* n_nt_entries,
                                const uint32_t* nt_allele, const uint32_t* nt_n_names, const char* names_blob,
                                uint64_t blob_len, uint64_t plane_cap_words, uint32_t* planes, uint8_t* site_flags,
                                uint8_t* bad_site, uint32_t* n_reads_out, uint32_t* row_words_out) {
  if ((n_sites && (!site_type || !n_depth_entries || !n_nt_entries || !site_flags)) || !n_reads_out || !row_words_out)
    return LGMI_ERR_ARG;
  // ---- pass 1: intern the read names in order of first appearance; (site, read) -> last allele wins
  // Open-addressing table of the distinct names (keys are views into the blob): 64-bit FNV-1a, linear
  // probing, at most half full.  Ids are handed out in order of first appearance.
  struct NameSlot {
    uint64_t hash;
    const char* ptr;
    uint32_t len, id;
  };
  std::vector<NameSlot> table(1024, NameSlot{0, nullptr, 0, 0});
  uint32_t n_names = 0;
  auto grow = [&]() {
    std::vector<NameSlot> bigger(table.size() * 2, NameSlot{0, nullptr, 0, 0});
    const size_t mask = bigger.size() - 1;
    for (const NameSlot& e : table) {
      if (!e.ptr) continue;
      size_t k = (size_t)e.hash & mask;
      while (bigger[k].ptr) k = (k + 1) & mask;
      bigger[k] = e;
    }
    table.swap(bigger);
  };
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> site_reads(n_sites);  // (read, allele) in encounter order
  // per read: the last site it was seen at (+1) and its position in that site's list -- the per-site
  // "read already listed" test without a hash map per site
  std::vector<uint32_t> seen_site, seen_slot;
  uint64_t cursor = 0;
  size_t nt_k = 0;
  for (uint32_t s = 0; s < n_sites; ++s) {
    for (uint32_t e = 0; e < n_nt_entries[s]; ++e, ++nt_k) {
      const uint32_t allele = nt_allele[nt_k];
      site_reads[s].reserve(site_reads[s].size() + nt_n_names[nt_k]);
      for (uint32_t q = 0; q < nt_n_names[nt_k]; ++q) {
        if (cursor > blob_len) return LGMI_ERR_ARG;
        const char* b = names_blob + cursor;
        uint64_t h = 1469598103934665603ull;
        uint64_t len = 0;
        const uint64_t room = blob_len - cursor;
        while (len < room && b[len] != '\n') {
          h = (h ^ (unsigned char)b[len]) * 1099511628211ull;
          ++len;
        }
        cursor += len + 1;
        h ^= h >> 29;  // the low bits index the table
        size_t k = (size_t)h & (table.size() - 1);
        uint32_t r;
        for (;;) {
          NameSlot& slot = table[k];
          if (!slot.ptr) {
            r = n_names++;
            slot = NameSlot{h, b, (uint32_t)len, r};
            seen_site.push_back(0u);
            seen_slot.push_back(0u);
            if ((size_t)n_names * 2 > table.size()) grow();
            break;
          }
          if (slot.hash == h && slot.len == len && memcmp(slot.ptr, b, len) == 0) {
            r = slot.id;
            break;
          }
          k = (k + 1) & (table.size() - 1);
        }
        if (seen_site[r] != s + 1u) {
          seen_site[r] = s + 1u;
          seen_slot[r] = (uint32_t)site_reads[s].size();
          site_reads[s].emplace_back(r, allele);
        } else {
          site_reads[s][seen_slot[r]].second = allele;  // mutual_information.py:15-16: dict() keeps the last
        }
      }
    }
  }
  const uint32_t R = n_names;
  const uint32_t W = 4u * ((R + 127u) / 128u);
  *n_reads_out = R;
  *row_words_out = W;
  if ((uint64_t)3u * n_sites * W > plane_cap_words) return LGMI_ERR_NOMEM;
  if (n_sites && W && !planes) return LGMI_ERR_ARG;
  if (n_sites && W) memset(planes, 0, (size_t)3u * n_sites * W * sizeof(uint32_t));
  // ---- pass 2: major / minor by depth (stable: ties keep the depth dict's order), labels -> planes
  size_t dk = 0;
  for (uint32_t s = 0; s < n_sites; ++s) {
    const uint32_t nd = n_depth_entries[s];
    uint32_t major = 0xffffffffu, minor = 0xffffffffu;
    int64_t d_major = 0, d_minor = 0;
    bool have_major = false, have_minor = false;
    for (uint32_t e = 0; e < nd; ++e, ++dk) {  // first strictly-greater wins == stable descending sort
      const int64_t d = depth_value[dk];
      if (!have_major || d > d_major) {
        if (have_major) {
          minor = major;
          d_minor = d_major;
          have_minor = true;
        }
        major = depth_allele[dk];
        d_major = d;
        have_major = true;
      } else if (!have_minor || d > d_minor) {
        minor = depth_allele[dk];
        d_minor = d;
        have_minor = true;
      }
    }
    if (bad_site) bad_site[s] = nd < 2 ? 1 : 0;  // the reference raises IndexError at :30 / :32 for such a site
    uint32_t* row = planes + (size_t)s * 3u * W;
    bool any_other = false;
    for (const auto& ra : site_reads[s]) {
      const uint32_t r = ra.first, w = r >> 5, bit = 1u << (r & 31u);
      row[2u * W + w] |= bit;                                      // covered
      if (have_major && ra.second == major) row[w] |= bit;         // label 2
      else if (have_minor && ra.second == minor) row[W + w] |= bit;  // label 1
      else any_other = true;                                       // label 0
    }
    site_flags[s] = (uint8_t)((site_type[s] & LGMI_SITE_TYPE_MASK) | (any_other ? LGMI_SITE_HAS_OTHER : 0u));
  }
  return LGMI_OK;
}
