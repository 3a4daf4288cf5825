// TEST INFRASTRUCTURE ONLY (oracle/): see CompactNSearch.h in this directory.  Compile this TU with
// -ffp-contract=off: the neighbour predicate must not be contracted into an FMA.
#include "CompactNSearch.h"
#include <algorithm>
#include <numeric>
#include <limits>
#include <stdexcept>
#include <cmath>

namespace CompactNSearch
{
NeighborhoodSearch::NeighborhoodSearch(Real r, bool) { set_radius(r); }

unsigned int NeighborhoodSearch::add_point_set(Real const* x, std::size_t n, bool is_dynamic, bool search_neighbors, bool find_neighbors, void* user_data)
{
	m_point_sets.push_back(PointSet(x, n, is_dynamic, user_data));
	const std::size_t size = m_table.size();
	for (std::size_t i = 0; i < size; ++i) m_table[i].push_back(find_neighbors ? 1 : 0);
	m_table.push_back(std::vector<unsigned char>(size + 1, search_neighbors ? 1 : 0));
	return static_cast<unsigned int>(m_point_sets.size() - 1);
}

void NeighborhoodSearch::resize_point_set(unsigned int i, Real const* x, std::size_t n)
{
	m_point_sets[i].m_x = x;
	m_point_sets[i].m_n = n;
	m_point_sets[i].m_sort_table.clear();
}

void NeighborhoodSearch::set_active(unsigned int index, bool search_neighbors, bool find_neighbors)
{
	const std::size_t size = m_table.size();
	for (std::size_t j = 0; j < size; ++j)
	{
		m_table[index][j] = search_neighbors ? 1 : 0;
		m_table[j][index] = find_neighbors ? 1 : 0;
	}
	m_table[index][index] = (search_neighbors && find_neighbors) ? 1 : 0;
}

void NeighborhoodSearch::set_active(bool active)
{
	for (auto& row : m_table) std::fill(row.begin(), row.end(), active ? 1 : 0);
}

void NeighborhoodSearch::cell_of(Real const* x, int c[3]) const
{
	for (int k = 0; k < 3; ++k)
	{
		if (x[k] >= 0.0) c[k] = static_cast<int>(m_inv_cell_size * x[k]);
		else c[k] = static_cast<int>(m_inv_cell_size * x[k]) - 1;
	}
}

static inline uint64_t spread3(uint64_t v)
{
	v &= 0x1fffffULL;
	v = (v | (v << 32)) & 0x1f00000000ffffULL;
	v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
	v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
	v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
	v = (v | (v << 2)) & 0x1249249249249249ULL;
	return v;
}

void NeighborhoodSearch::z_sort()
{
	for (PointSet& d : m_point_sets)
	{
		const std::size_t n = d.n_points();
		std::vector<uint64_t> code(n);
		#pragma omp parallel for schedule(static)
		for (long i = 0; i < (long)n; ++i)
		{
			int c[3];
			cell_of(d.point((unsigned int)i), c);
			// shift to non-negative (upstream subtracts INT_MIN+1; 2^20 keeps 21 bits/axis)
			code[i] = spread3((uint64_t)(c[0] + (1 << 20))) | (spread3((uint64_t)(c[1] + (1 << 20))) << 1) | (spread3((uint64_t)(c[2] + (1 << 20))) << 2);
		}
		d.m_sort_table.resize(n);
		std::iota(d.m_sort_table.begin(), d.m_sort_table.end(), 0u);
		std::stable_sort(d.m_sort_table.begin(), d.m_sort_table.end(), [&](unsigned int a, unsigned int b) { return code[a] < code[b]; });
	}
}

void NeighborhoodSearch::find_neighbors(bool)
{
	const unsigned int nsets = (unsigned int)m_point_sets.size();
	// which sets take part
	std::vector<char> searches(nsets, 0), found(nsets, 0);
	for (unsigned int a = 0; a < nsets; ++a)
		for (unsigned int b = 0; b < nsets; ++b)
			if (m_table[a][b] && m_point_sets[a].n_points() > 0 && m_point_sets[b].n_points() > 0) { searches[a] = 1; found[b] = 1; }

	// bounding box in cell coordinates over all participating sets
	int lo[3] = { std::numeric_limits<int>::max(), std::numeric_limits<int>::max(), std::numeric_limits<int>::max() };
	int hi[3] = { std::numeric_limits<int>::lowest(), std::numeric_limits<int>::lowest(), std::numeric_limits<int>::lowest() };
	bool any = false;
	for (unsigned int s = 0; s < nsets; ++s)
	{
		if (!searches[s] && !found[s]) continue;
		const PointSet& d = m_point_sets[s];
		const long n = (long)d.n_points();
		if (n > 0) any = true;
		#pragma omp parallel
		{
			int tlo[3] = { std::numeric_limits<int>::max(), std::numeric_limits<int>::max(), std::numeric_limits<int>::max() };
			int thi[3] = { std::numeric_limits<int>::lowest(), std::numeric_limits<int>::lowest(), std::numeric_limits<int>::lowest() };
			#pragma omp for schedule(static) nowait
			for (long i = 0; i < n; ++i)
			{
				int c[3];
				cell_of(d.point((unsigned int)i), c);
				for (int k = 0; k < 3; ++k) { tlo[k] = std::min(tlo[k], c[k]); thi[k] = std::max(thi[k], c[k]); }
			}
			#pragma omp critical
			for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], tlo[k]); hi[k] = std::max(hi[k], thi[k]); }
		}
	}
	for (unsigned int a = 0; a < nsets; ++a)
	{
		PointSet& d = m_point_sets[a];
		d.m_neighbors.resize(nsets);
		for (unsigned int b = 0; b < nsets; ++b)
		{
			if (m_table[a][b]) { d.m_neighbors[b].resize(d.n_points()); }
			else { for (auto& l : d.m_neighbors[b]) l.clear(); d.m_neighbors[b].resize(d.n_points()); }
		}
	}
	if (!any) return;

	const long long nx = (long long)hi[0] - lo[0] + 1, ny = (long long)hi[1] - lo[1] + 1, nz = (long long)hi[2] - lo[2] + 1;
	const long long ncells = nx * ny * nz;
	if (ncells > 2000000000LL) throw std::runtime_error("CompactNSearch stand-in: grid too large");

	// counting sort of each found set into cells
	std::vector<std::vector<unsigned int>> cell_start(nsets), sorted(nsets);
	for (unsigned int s = 0; s < nsets; ++s)
	{
		if (!found[s]) continue;
		const PointSet& d = m_point_sets[s];
		const std::size_t n = d.n_points();
		std::vector<unsigned int> cid(n);
		cell_start[s].assign((std::size_t)ncells + 1, 0u);
		// parallel counting sort; every cell segment ends up in ascending point order (= the order a serial pass produces)
		unsigned int* cnt = cell_start[s].data() + 1;
		#pragma omp parallel for schedule(static)
		for (long i = 0; i < (long)n; ++i)
		{
			int c[3];
			cell_of(d.point((unsigned int)i), c);
			cid[i] = (unsigned int)(((long long)(c[0] - lo[0]) * ny + (c[1] - lo[1])) * nz + (c[2] - lo[2]));
			#pragma omp atomic
			cnt[cid[i]]++;
		}
		for (long long c = 0; c < ncells; ++c) cell_start[s][c + 1] += cell_start[s][c];
		sorted[s].resize(n);
		std::vector<unsigned int> cursor(cell_start[s].begin(), cell_start[s].end() - 1);
		#pragma omp parallel for schedule(static)
		for (long i = 0; i < (long)n; ++i)
		{
			unsigned int slot;
			#pragma omp atomic capture
			slot = cursor[cid[i]]++;
			sorted[s][slot] = (unsigned int)i;
		}
		#pragma omp parallel for schedule(static, 4096)
		for (long long c = 0; c < ncells; ++c)
		{
			const unsigned int k0 = cell_start[s][c], k1 = cell_start[s][c + 1];
			if (k1 - k0 > 1) std::sort(sorted[s].begin() + k0, sorted[s].begin() + k1);
		}
	}

	const Real r2 = m_r2;
	for (unsigned int a = 0; a < nsets; ++a)
	{
		if (!searches[a]) continue;
		PointSet& da = m_point_sets[a];
		const long n = (long)da.n_points();
		#pragma omp parallel for schedule(dynamic, 256)
		for (long i = 0; i < n; ++i)
		{
			Real const* xa = da.point((unsigned int)i);
			int c[3];
			cell_of(xa, c);
			for (unsigned int b = 0; b < nsets; ++b)
			{
				if (!m_table[a][b] || !found[b]) continue;
				std::vector<unsigned int>& out = da.m_neighbors[b][i];
				out.clear();
				const PointSet& db = m_point_sets[b];
				for (int dx = -1; dx <= 1; ++dx)
				{
					const long long cx = (long long)c[0] + dx - lo[0];
					if (cx < 0 || cx >= nx) continue;
					for (int dy = -1; dy <= 1; ++dy)
					{
						const long long cy = (long long)c[1] + dy - lo[1];
						if (cy < 0 || cy >= ny) continue;
						// the three z-cells are contiguous in the linear index
						const long long cz0 = std::max<long long>((long long)c[2] - 1 - lo[2], 0), cz1 = std::min<long long>((long long)c[2] + 1 - lo[2], nz - 1);
						if (cz0 > cz1) continue;
						const std::size_t base = (std::size_t)((cx * ny + cy) * nz);
						const unsigned int k0 = cell_start[b][base + cz0], k1 = cell_start[b][base + cz1 + 1];
						for (unsigned int k = k0; k < k1; ++k)
						{
							const unsigned int j = sorted[b][k];
							if (a == b && j == (unsigned int)i) continue;
							Real const* xb = db.point(j);
							Real tmp = xa[0] - xb[0];
							Real l2 = tmp * tmp;
							tmp = xa[1] - xb[1];
							l2 += tmp * tmp;
							tmp = xa[2] - xb[2];
							l2 += tmp * tmp;
							if (l2 < r2) out.push_back(j);
						}
					}
				}
			}
		}
	}
}
}
