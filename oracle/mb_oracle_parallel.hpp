// TEST INFRASTRUCTURE ONLY -- CPU oracle, shared-memory chunk exchange (see mb_oracle.hpp header for the rules).
// Restates /root/reference/src/parallel.jl; line numbers cite that file.  This is the reference's only
// "distributed" component; the B200 path replaces it with a slab partition + NCCL neighbour exchange whose
// post-exchange per-cell logical order ([own | from chunk 1 | from chunk 2 ...]) is the same.
#pragma once
#include <utility>

#include "mb_oracle.hpp"

namespace mbo {

struct CellChunk { int64_t lo, hi; };  // contiguous 1-based inclusive cell range (ChunkSplitters.chunks(1:nx; n))

struct ChunkExchanger {  // :21-46
    int64_t n_chunks, n_cells;
    std::vector<ParticleIndexer> indexer;  // [chunk + n_chunks*cell], 0-based storage
    ChunkExchanger(int64_t nch, int64_t nc) : n_chunks(nch), n_cells(nc), indexer(nch * nc) {}
    ParticleIndexer& at(int64_t chunk, int64_t cell) { return indexer[(chunk - 1) + n_chunks * (cell - 1)]; }
};
inline void reset_exchanger(ChunkExchanger& ex, int64_t chunk_id) {  // :57-67
    for (int64_t i = 1; i <= ex.n_cells; i++) {
        ParticleIndexer& q = ex.at(chunk_id, i);
        q.n_group1 = 0; q.start1 = 0; q.end1 = -1; q.n_group2 = 0; q.start2 = 0; q.end2 = -1;
    }
}

// :95-172
inline void push_particles(ChunkExchanger& ex, std::vector<ParticleVector*>& pvs, std::vector<ParticleIndexerArray*>& pias, int64_t species,
                           int64_t i, int64_t j, int64_t offset_ij, int64_t s_ci_ij2, int64_t e_ci_ij) {
    ParticleVector& Pi = *pvs[i - 1];
    ParticleVector& Pj = *pvs[j - 1];
    ParticleIndexerArray& Ai = *pias[i - 1];
    ParticleIndexerArray& Aj = *pias[j - 1];
    auto move_range = [&](int64_t cell, int64_t n_move, int64_t extra_offset) {
        ParticleIndexer& q = ex.at(i, cell);
        q.start2 = Aj.n_total[species - 1] + 1;
        Aj.n_total[species - 1] += n_move;
        q.end2 = Aj.n_total[species - 1];
        q.n_group2 = n_move;
        const int64_t s2 = q.start2, e2 = q.end2;
        const int64_t offset = -s2 + Ai.at(cell, species).start1 + extra_offset;
        for (int64_t pid = s2; pid <= e2; pid++) {
            update_particle_buffer_new_particle(Pj, pid);
            Pj[pid] = Pi[pid + offset];
            Pi.nbuffer += 1;
            Pi.buffer[Pi.nbuffer - 1] = Pi.index[pid + offset - 1];
        }
        ParticleIndexer& a = Ai.at(cell, species);
        a.n_local = 0; a.n_group1 = 0; a.start1 = 0; a.end1 = -1;
    };
    int64_t cell = s_ci_ij2;
    if (Ai.at(cell, species).n_group1 > 0 && offset_ij < Ai.at(cell, species).n_group1)
        move_range(cell, Ai.at(cell, species).n_group1 - offset_ij, offset_ij);
    for (cell = s_ci_ij2 + 1; cell <= e_ci_ij; cell++)
        if (Ai.at(cell, species).n_group1 > 0) move_range(cell, Ai.at(cell, species).n_group1, 0);
}

// :199-249 ; returns (s_ci_ij2, offset_ij)
inline std::pair<int64_t, int64_t> update_swap_indexing(ChunkExchanger& ex, std::vector<ParticleIndexerArray*>& pias, int64_t species, int64_t i,
                                                        int64_t /*j*/, int64_t s_ci_ij, int64_t e_ci_ij, int64_t s_ji, int64_t n_swap) {
    int64_t s_ci_ij2 = s_ci_ij, offset_ij = 0;
    ParticleIndexerArray& Ai = *pias[i - 1];
    for (int64_t cell = s_ci_ij; cell <= e_ci_ij; cell++) {
        ParticleIndexer& a = Ai.at(cell, species);
        const int64_t offset = std::min(a.n_group1, n_swap);
        if (offset == a.n_group1) { a.n_local = 0; a.n_group1 = 0; a.start1 = 0; a.end1 = -1; }
        if (offset > 0) {
            ParticleIndexer& q = ex.at(i, cell);
            q.start1 = s_ji;
            s_ji += offset - 1;
            q.end1 = s_ji;
            q.n_group1 = offset;
            n_swap -= offset;
            s_ci_ij2 = cell;
            offset_ij = offset;
            if (n_swap <= 0) break;
            s_ji += 1;
        }
    }
    return {s_ci_ij2, offset_ij};
}

// :281-414
inline void exchange_particles(ChunkExchanger& ex, std::vector<ParticleVector*>& pvs, std::vector<ParticleIndexerArray*>& pias,
                               const std::vector<CellChunk>& chunks, int64_t species, int64_t i, int64_t j) {
    ParticleIndexerArray& Ai = *pias[i - 1];
    ParticleIndexerArray& Aj = *pias[j - 1];
    auto first_last = [&](ParticleIndexerArray& A, const CellChunk& ch, int64_t& s, int64_t& sc, int64_t& e, int64_t& ec) {
        s = 0; sc = 0; e = -1; ec = 0;
        for (int64_t c = ch.lo; c <= ch.hi; c++)
            if (A.at(c, species).start1 > 0) { s = A.at(c, species).start1; sc = c; break; }
        for (int64_t c = ch.hi; c >= ch.lo; c--)
            if (A.at(c, species).end1 > 0) { e = A.at(c, species).end1; ec = c; break; }
    };
    int64_t s_ij, s_ci_ij, e_ij, e_ci_ij, s_ji, s_ci_ji, e_ji, e_ci_ji;
    first_last(Ai, chunks[j - 1], s_ij, s_ci_ij, e_ij, e_ci_ij);
    int64_t np_i_to_j = e_ij - s_ij + 1;
    first_last(Aj, chunks[i - 1], s_ji, s_ci_ji, e_ji, e_ci_ji);
    int64_t np_j_to_i = e_ji - s_ji + 1;

    int64_t inc_i = np_j_to_i > 0 ? np_j_to_i : 0;
    inc_i = np_i_to_j > 0 ? inc_i - np_i_to_j : inc_i;
    int64_t inc_j = np_i_to_j > 0 ? np_i_to_j : 0;
    inc_j = np_j_to_i > 0 ? inc_j - np_j_to_i : inc_j;
    ParticleVector& Pi = *pvs[i - 1];
    ParticleVector& Pj = *pvs[j - 1];
    if (Pi.length() < Ai.n_total[species - 1] + inc_i) Pi.resize(Pi.length() + inc_i + DELTA_PARTICLES);
    if (Pj.length() < Aj.n_total[species - 1] + inc_j) Pj.resize(Pj.length() + inc_j + DELTA_PARTICLES);

    const int64_t n_swap = std::min(np_i_to_j, np_j_to_i);
    int64_t offset_ij = 0, offset_ji = 0, s_ci_ij2 = s_ci_ij, s_ci_ji2 = s_ci_ji;
    if (n_swap > 0) {
        for (int64_t nsw = 1; nsw <= n_swap; nsw++) swap_particles(Pi, Pj, s_ij + nsw - 1, s_ji + nsw - 1);
        auto r1 = update_swap_indexing(ex, pias, species, i, j, s_ci_ij, e_ci_ij, s_ji, n_swap);
        s_ci_ij2 = r1.first; offset_ij = r1.second;
        auto r2 = update_swap_indexing(ex, pias, species, j, i, s_ci_ji, e_ci_ji, s_ij, n_swap);
        s_ci_ji2 = r2.first; offset_ji = r2.second;
        np_i_to_j -= n_swap;
        np_j_to_i -= n_swap;
    }
    if (np_i_to_j > 0) push_particles(ex, pvs, pias, species, i, j, offset_ij, s_ci_ij2, e_ci_ij);
    else if (np_j_to_i > 0) push_particles(ex, pvs, pias, species, j, i, offset_ji, s_ci_ji2, e_ci_ji);
}
// :443-450
inline void exchange_particles_all(ChunkExchanger& ex, std::vector<ParticleVector*>& pvs, std::vector<ParticleIndexerArray*>& pias,
                                   const std::vector<CellChunk>& chunks, int64_t species) {
    const int64_t n = (int64_t)chunks.size();
    for (int64_t i = 1; i <= n - 1; i++)
        for (int64_t j = i + 1; j <= n; j++) exchange_particles(ex, pvs, pias, chunks, species, i, j);
}

// :467-532
inline void sort_particles_after_exchange(ChunkExchanger& ex, GridSortInPlace& gs, ParticleVector& pv, ParticleIndexerArray& pia,
                                          const CellChunk& chunk, int64_t species) {
    int64_t n_tot = pia.n_total[species - 1];
    if (n_tot > (int64_t)gs.sorted_indices.size()) gs.sorted_indices.resize(n_tot + DELTA_PARTICLES);
    int64_t ci = 0, offset = 0;
    n_tot = 0;
    for (int64_t cell = chunk.lo; cell <= chunk.hi; cell++) {
        ParticleIndexer& a = pia.at(cell, species);
        gs.cell_counts[cell - 1] = a.n_group1;
        for (int64_t i = a.start1; i <= a.end1; i++) gs.sorted_indices[ci++] = pv.index[i - 1];
        for (int64_t ch = 1; ch <= ex.n_chunks; ch++) {
            const ParticleIndexer& q = ex.at(ch, cell);
            gs.cell_counts[cell - 1] += q.n_group1;
            gs.cell_counts[cell - 1] += q.n_group2;
            for (int64_t i = q.start1; i <= q.end1; i++) gs.sorted_indices[ci++] = pv.index[i - 1];
            for (int64_t i = q.start2; i <= q.end2; i++) gs.sorted_indices[ci++] = pv.index[i - 1];
        }
        n_tot += gs.cell_counts[cell - 1];
        a.n_group1 = gs.cell_counts[cell - 1];
        a.n_local = gs.cell_counts[cell - 1];
        if (a.n_group1 > 0) { a.start1 = offset + 1; a.end1 = offset + gs.cell_counts[cell - 1]; }
        else { a.start1 = 0; a.end1 = -1; }
        offset += gs.cell_counts[cell - 1];
        a.n_group2 = 0; a.start2 = 0; a.end2 = -1;
    }
    pia.n_total[species - 1] = n_tot;
    for (int64_t i = 1; i <= n_tot; i++) pv.index[i - 1] = gs.sorted_indices[i - 1];
}

// :559-581
inline std::vector<std::vector<std::pair<int64_t, int64_t>>> generate_1_factorization(int64_t N) {
    std::vector<std::vector<std::pair<int64_t, int64_t>>> lol;
    for (int64_t i = 1; i <= N; i++)
        for (int64_t j = i + 1; j <= N; j++) {
            bool assigned = false;
            for (auto& inner : lol) {
                bool used = false;
                for (auto& p : inner)
                    if (p.first == i || p.second == i || p.first == j || p.second == j) { used = true; break; }
                if (!used) { inner.push_back({i, j}); assigned = true; break; }
            }
            if (!assigned) lol.push_back({{i, j}});
        }
    return lol;
}

}  // namespace mbo
