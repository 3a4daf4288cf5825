// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for CompactNSearch
// (InteractiveComputerGraphics/CompactNSearch @ a9ab7c71ce264487660ecbaf81b5060bda462722, pinned in the
// reference at CMake/NeighborhoodSearch.cmake:35-36), fetched by the reference at build time and absent from
// /root/reference.  Written from scratch.  API = exactly what the reference calls (Simulation.cpp:149-151,617,626,
// 698-755; Simulation.h:354,359,451-473; FluidModel.cpp:214-216,322-346; BoundaryModel_Akinci2012.cpp:52-69,
// 108-126; SimulationDataDFSPH.cpp:94-98).  Behavioural contract restated from the published algorithm:
//   * neighbour predicate: l2 = dx*dx; l2 += dy*dy; l2 += dz*dz (each op rounded in Real, no FMA);  j is a
//     neighbour of i  iff  l2 < r*r  (strict) and (set_a,i) != (set_b,j);
//   * cell size = r; cell index per axis = (int)(x/r) for x >= 0 and (int)(x/r) - 1 for x < 0;
//   * z_sort(): permutation of each point set by the Morton code of the cell index; sort_field() applies it;
//   * activation table semantics of add_point_set / set_active.
// Unlike the upstream library (spatial hashing + per-point spin locks, thread-dependent list order) the lists
// produced here are deterministic (cell-major, ascending index inside a cell).  PARITY NOTE: neighbour-set parity
// against upstream CompactNSearch itself is UNPINNED (the reference ships no neighbour-search test vectors).
#pragma once
#include <vector>
#include <cstddef>
#include <cstdint>

namespace CompactNSearch
{
#ifdef USE_DOUBLE
	using Real = double;
#else
	using Real = float;
#endif

	class NeighborhoodSearch;

	class PointSet
	{
	public:
		std::size_t n_neighbors(unsigned int point_set, unsigned int i) const { return m_neighbors[point_set][i].size(); }
		unsigned int neighbor(unsigned int point_set, unsigned int i, unsigned int k) const { return m_neighbors[point_set][i][k]; }
		std::vector<unsigned int> const& neighbor_list(unsigned int point_set, unsigned int i) const { return m_neighbors[point_set][i]; }
		std::size_t n_points() const { return m_n; }
		bool is_dynamic() const { return m_dynamic; }
		void set_dynamic(bool v) { m_dynamic = v; }
		void* get_user_data() { return m_user_data; }
		void* get_user_data() const { return m_user_data; }
		Real const* GetPoints() const { return m_x; }
		Real const* point(unsigned int i) const { return &m_x[3 * i]; }

		/** Reorders an array according to the permutation computed by the last z_sort(). */
		template <typename T>
		void sort_field(T* lst) const
		{
			if (m_sort_table.empty()) return;
			std::vector<T> tmp(lst, lst + m_sort_table.size());
			for (std::size_t i = 0; i < m_sort_table.size(); ++i) lst[i] = tmp[m_sort_table[i]];
		}

	private:
		friend class NeighborhoodSearch;
		PointSet(Real const* x, std::size_t n, bool dynamic, void* user_data) : m_x(x), m_n(n), m_dynamic(dynamic), m_user_data(user_data) {}
		Real const* m_x;
		std::size_t m_n;
		bool m_dynamic;
		void* m_user_data;
		std::vector<unsigned int> m_sort_table;
		std::vector<std::vector<std::vector<unsigned int>>> m_neighbors;   // [other set][point] -> list
	};

	class NeighborhoodSearch
	{
	public:
		NeighborhoodSearch(Real r, bool erase_empty_cells = false);
		virtual ~NeighborhoodSearch() {}

		PointSet const& point_set(unsigned int i) const { return m_point_sets[i]; }
		PointSet& point_set(unsigned int i) { return m_point_sets[i]; }
		std::size_t n_point_sets() const { return m_point_sets.size(); }
		std::vector<PointSet> const& point_sets() const { return m_point_sets; }
		std::vector<PointSet>& point_sets() { return m_point_sets; }

		unsigned int add_point_set(Real const* x, std::size_t n, bool is_dynamic = true, bool search_neighbors = true, bool find_neighbors = true, void* user_data = nullptr);
		void resize_point_set(unsigned int i, Real const* x, std::size_t n);
		void find_neighbors(bool points_changed = true);
		void update_point_sets() {}
		void z_sort();
		void reset() {}

		Real radius() const { return m_r; }
		void set_radius(Real r) { m_r = r; m_r2 = r * r; m_inv_cell_size = static_cast<Real>(1.0 / r); }

		void set_active(unsigned int i, unsigned int j, bool active) { m_table[i][j] = active ? 1 : 0; }
		void set_active(unsigned int i, bool search_neighbors = true, bool find_neighbors = true);
		void set_active(bool active);
		bool is_active(unsigned int i, unsigned int j) const { return m_table[i][j] != 0; }

	private:
		void cell_of(Real const* x, int c[3]) const;
		Real m_r, m_r2, m_inv_cell_size;
		std::vector<PointSet> m_point_sets;
		std::vector<std::vector<unsigned char>> m_table;   // [searching set][found set]
	};
}
