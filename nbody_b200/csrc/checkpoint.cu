// Checkpoint / restart and the variable-time-step rule.
//
// The reference can only append positions to particles.csv (src/main.cpp:88-95): velocities,
// masses and charges are dropped, so a run cannot be resumed. A checkpoint here is the full
// state the C ABI exposes — configuration, time accumulator, step count, next time step, the
// 48-byte particle records in particles() order and the permutation — in one flat file
// (layout in include/nbody_cuda.h). Reading, writing and validating a file is host arithmetic
// and needs no device; only save (device -> file) and load (file -> new simulation) touch CUDA.
#include <cerrno>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include <unistd.h>

#include "common.cuh"

namespace nbody {

namespace {

struct FileCloser { void operator()(std::FILE* f) const { if (f) std::fclose(f); } };
using File = std::unique_ptr<std::FILE, FileCloser>;

// sum_i (w_i + golden) * (2 i + 1) mod 2^64 over the 64-bit little-endian words of `bytes` (zero-padded), continuing at word `i0`
uint64_t word_sum(const void* data, size_t bytes, uint64_t& i0) {
	const unsigned char* p = static_cast<const unsigned char*>(data);
	uint64_t acc = 0;
	size_t k = 0;
	for (; k + 8 <= bytes; k += 8) {
		uint64_t w;
		std::memcpy(&w, p + k, 8);
		acc += (w + 0x9E3779B97F4A7C15ull) * (2 * i0 + 1);
		++i0;
	}
	if (k < bytes) {
		uint64_t w = 0;
		std::memcpy(&w, p + k, bytes - k);
		acc += (w + 0x9E3779B97F4A7C15ull) * (2 * i0 + 1);
		++i0;
	}
	return acc;
}

// Version 2 files: the sum starts with the header itself (its checksum field read as zero), so a changed time, step count,
// next step or configuration is detected like a changed particle; version 1 files (payload only) are still read.
// (Works on the bytes of the very object that is written to / was read from the file, tail padding included: a by-value copy of
// the struct need not preserve padding bytes.)
uint64_t file_checksum(const nbody_checkpoint_header& h, const nbody_particle* particles, const uint32_t* orig, uint64_t n) {
	uint64_t i = 0, c = 0;
	if (h.version >= 2u) {
		unsigned char raw[sizeof(nbody_checkpoint_header)];
		std::memcpy(raw, &h, sizeof(raw));
		std::memset(raw + offsetof(nbody_checkpoint_header, checksum), 0, sizeof(uint64_t));
		c = word_sum(raw, sizeof(raw), i);
	}
	c += word_sum(particles, (size_t) n * sizeof(nbody_particle), i);
	c += word_sum(orig, (size_t) n * sizeof(uint32_t), i);
	return c;
}

int fail_io(const std::string& what, const char* path) {
	set_error(what + " '" + (path ? path : "(null)") + "': " + std::strerror(errno));
	return NBODY_ERR_INVALID;
}

int read_header(std::FILE* f, const char* path, nbody_checkpoint_header* h) {
	if (std::fread(h, sizeof(*h), 1, f) != 1) { set_error(std::string("checkpoint '") + path + "' is shorter than its header"); return NBODY_ERR_INVALID; }
	if (h->magic != NBODY_CHECKPOINT_MAGIC) { set_error(std::string("'") + path + "' is not an nbody checkpoint (bad magic)"); return NBODY_ERR_INVALID; }
	if (h->version < 1u || h->version > NBODY_CHECKPOINT_VERSION || h->header_bytes != sizeof(nbody_checkpoint_header)) {
		set_error("checkpoint version / header size not understood by this library");
		return NBODY_ERR_INVALID;
	}
	if (h->config.abi_version != NBODY_CUDA_ABI_VERSION) { set_error("checkpoint was written with a different ABI version"); return NBODY_ERR_INVALID; }
	if (h->n_particles == 0 || h->n_particles > 0xfffffff0ull) { set_error("checkpoint particle count out of range"); return NBODY_ERR_INVALID; }
	if (std::fseek(f, 0, SEEK_END) != 0) return fail_io("cannot seek in", path);
	const long long size = std::ftell(f);
	const unsigned long long want = sizeof(*h) + h->n_particles * (sizeof(nbody_particle) + sizeof(uint32_t));
	if (size < 0 || (unsigned long long) size != want) { set_error("checkpoint is truncated or has trailing bytes (size does not match its header)"); return NBODY_ERR_INVALID; }
	if (std::fseek(f, (long) sizeof(*h), SEEK_SET) != 0) return fail_io("cannot seek in", path);
	return NBODY_OK;
}

int write_file(const char* path, nbody_checkpoint_header h, const nbody_particle* particles, const uint32_t* orig) {
	const uint64_t n = h.n_particles;
	std::vector<uint32_t> ident;
	if (!orig) {
		ident.resize(n);
		for (uint64_t i = 0; i < n; ++i) ident[i] = (uint32_t) i;
		orig = ident.data();
	}
	h.magic = NBODY_CHECKPOINT_MAGIC;
	h.version = NBODY_CHECKPOINT_VERSION;
	h.header_bytes = (uint32_t) sizeof(h);
	{  // the padding behind the last field is written as zeros, whatever the caller's struct held there
		constexpr size_t used = offsetof(nbody_checkpoint_header, config) + sizeof(nbody_cuda_config);
		std::memset(reinterpret_cast<unsigned char*>(&h) + used, 0, sizeof(h) - used);
	}
	h.checksum = file_checksum(h, particles, orig, n);
	// write next to the target and rename, so that an interrupted save never leaves a half-written checkpoint behind
	const std::string tmp = std::string(path) + ".partial";
	File f(std::fopen(tmp.c_str(), "wb"));
	if (!f) return fail_io("cannot create", tmp.c_str());
	const bool ok = std::fwrite(&h, sizeof(h), 1, f.get()) == 1 && std::fwrite(particles, sizeof(nbody_particle), n, f.get()) == n &&
	                std::fwrite(orig, sizeof(uint32_t), n, f.get()) == n && std::fflush(f.get()) == 0 &&
	                fsync(fileno(f.get())) == 0;  // on disk before the rename makes it the checkpoint
	const int err = errno;                         // (the cleanup below may overwrite it)
	f.reset();
	if (!ok) { std::remove(tmp.c_str()); errno = err; return fail_io("short write to", tmp.c_str()); }
	if (std::rename(tmp.c_str(), path) != 0) { const int e2 = errno; std::remove(tmp.c_str()); errno = e2; return fail_io("cannot rename the finished checkpoint to", path); }
	return NBODY_OK;
}

}  // namespace

float next_time_step(const nbody_cuda_config& cfg, float acc_max) {
	if (!(cfg.time_step_eta > 0.0f) || !(acc_max > 0.0f) || !std::isfinite(acc_max)) return cfg.time_step;
	const float len = cfg.softening > 0.0f ? cfg.softening : std::ldexp(cfg.bounds[0], -(int) cfg.max_depth);
	float dt = cfg.time_step_eta * std::sqrt(len / acc_max);
	const float hi = cfg.time_step_max > 0.0f ? cfg.time_step_max : cfg.time_step;
	if (!(dt < hi)) dt = hi;
	if (cfg.time_step_min > 0.0f && dt < cfg.time_step_min) dt = cfg.time_step_min;
	return dt;
}

}  // namespace nbody

using namespace nbody;

extern "C" {

float nbody_cuda_next_time_step(const nbody_cuda_config* cfg, float acc_max) { return cfg ? next_time_step(*cfg, acc_max) : 0.0f; }

int nbody_cuda_set_time_step(nbody_cuda_sim* sim, float dt) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s) { set_error("NULL simulation"); return NBODY_ERR_INVALID; }
	if (!(dt > 0.0f) || !std::isfinite(dt)) { set_error("set_time_step: dt must be finite and > 0"); return NBODY_ERR_INVALID; }
	s->dt = dt;
	return NBODY_OK;
}

int nbody_cuda_get_time_step(nbody_cuda_sim* sim, float* next_dt, float* last_dt, float* acc_max) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s) { set_error("NULL simulation"); return NBODY_ERR_INVALID; }
	if (next_dt) *next_dt = s->dt;
	if (last_dt) *last_dt = s->dt_last;
	if (acc_max) *acc_max = s->acc_max;
	return NBODY_OK;
}

int nbody_cuda_get_time(nbody_cuda_sim* sim, float* time, uint64_t* steps_done) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s) { set_error("NULL simulation"); return NBODY_ERR_INVALID; }
	if (time) *time = s->time;
	if (steps_done) *steps_done = s->steps_base + s->steps_done;
	return NBODY_OK;
}

int nbody_cuda_checkpoint_info(const char* path, nbody_checkpoint_header* header) {
	if (!path || !header) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	File f(std::fopen(path, "rb"));
	if (!f) return fail_io("cannot open", path);
	return read_header(f.get(), path, header);
}

int nbody_cuda_checkpoint_read(const char* path, nbody_particle* particles, uint32_t* orig_index, uint64_t capacity) {
	if (!path) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	File f(std::fopen(path, "rb"));
	if (!f) return fail_io("cannot open", path);
	nbody_checkpoint_header h;
	int rc = read_header(f.get(), path, &h);
	if (rc) return rc;
	const uint64_t n = h.n_particles;
	if ((particles || orig_index) && capacity < n) { set_error("checkpoint_read: capacity is smaller than the checkpoint's particle count"); return NBODY_ERR_INVALID; }
	// the checksum covers both arrays, so both are read even when the caller wants one
	std::vector<nbody_particle> ptmp;
	std::vector<uint32_t> otmp;
	if (!particles) { ptmp.resize(n); particles = ptmp.data(); }
	if (!orig_index) { otmp.resize(n); orig_index = otmp.data(); }
	if (std::fread(particles, sizeof(nbody_particle), n, f.get()) != n || std::fread(orig_index, sizeof(uint32_t), n, f.get()) != n)
		return fail_io("short read from", path);
	if (file_checksum(h, particles, orig_index, n) != h.checksum) { set_error(std::string("checkpoint '") + path + "' is corrupt (checksum mismatch)"); return NBODY_ERR_INVALID; }
	return NBODY_OK;
}

int nbody_cuda_checkpoint_write(const char* path, const nbody_checkpoint_header* header, const nbody_particle* particles,
                                const uint32_t* orig_index) {
	if (!path || !header || !particles) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	if (header->n_particles == 0 || header->n_particles > 0xfffffff0ull) { set_error("checkpoint_write: particle count out of range"); return NBODY_ERR_INVALID; }
	if (header->config.abi_version != NBODY_CUDA_ABI_VERSION) { set_error("checkpoint_write: header->config.abi_version mismatch"); return NBODY_ERR_INVALID; }
	return write_file(path, *header, particles, orig_index);
}

int nbody_cuda_checkpoint_save(nbody_cuda_sim* sim, const char* path) {
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (!s || !path) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	if (s->let) {
		set_error("checkpoint_save: a rank of a partitioned run holds only its own particles; save them per rank with "
		          "nbody_cuda_get_owned_particles + nbody_cuda_get_permutation (the checkpoint format is single-rank)");
		return NBODY_ERR_STATE;
	}
	std::vector<nbody_particle> particles(s->n);
	std::vector<uint32_t> orig(s->n);
	int rc = nbody_cuda_get_particles(sim, particles.data(), s->n);
	if (rc) return rc;
	if ((rc = nbody_cuda_get_permutation(sim, orig.data(), s->n))) return rc;
	nbody_checkpoint_header h;
	std::memset(&h, 0, sizeof(h));
	h.n_particles = s->n; h.steps_done = s->steps_base + s->steps_done; h.time = s->time;
	h.next_time_step = s->dt; h.last_time_step = s->dt_last; h.last_acc_max = s->acc_max;
	h.config = s->cfg;
	return write_file(path, h, particles.data(), orig.data());
}

int nbody_cuda_checkpoint_load(const char* path, const nbody_cuda_config* cfg, nbody_cuda_sim** out) {
	if (!path || !out) { set_error("NULL argument"); return NBODY_ERR_INVALID; }
	*out = nullptr;
	nbody_checkpoint_header h;
	int rc = nbody_cuda_checkpoint_info(path, &h);
	if (rc) return rc;
	std::vector<nbody_particle> particles(h.n_particles);
	std::vector<uint32_t> orig(h.n_particles);
	if ((rc = nbody_cuda_checkpoint_read(path, particles.data(), orig.data(), h.n_particles))) return rc;
	{  // the stored identities must be a permutation of [0, n): callers index with them (P[inverse]), a foreign file must not get that far
		std::vector<uint8_t> seen(h.n_particles, 0);
		for (uint64_t i = 0; i < h.n_particles; ++i) {
			if (orig[i] >= h.n_particles || seen[orig[i]]) {
				set_error(std::string("checkpoint '") + path + "': the stored particle identities are not a permutation of [0, n)");
				return NBODY_ERR_INVALID;
			}
			seen[orig[i]] = 1;
		}
	}
	nbody_cuda_config use = cfg ? *cfg : h.config;
	if (!cfg) use.device = -1;  // the ordinal of the writing process means nothing here
	use.flags &= ~(uint32_t) NBODY_FLAG_PARTITIONED;  // (a single-rank object)
	nbody_cuda_sim* sim = nullptr;
	if ((rc = nbody_cuda_create(&use, particles.data(), h.n_particles, &sim))) return rc;
	Sim* s = reinterpret_cast<Sim*>(sim);
	if (cudaMemcpy(s->orig[0], orig.data(), h.n_particles * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
		set_error("checkpoint_load: upload of the permutation failed");
		nbody_cuda_destroy(sim);
		return NBODY_ERR_CUDA;
	}
	s->time = h.time; s->steps_base = h.steps_done;
	s->dt_last = h.last_time_step; s->acc_max = h.last_acc_max;
	// the stored step continues the sequence only under the rule that produced it; a new configuration starts from its own time_step
	s->dt = (cfg && (cfg->time_step != h.config.time_step || cfg->time_step_eta != h.config.time_step_eta)) ? use.time_step : h.next_time_step;
	if (!(s->dt > 0.0f)) s->dt = use.time_step;
	*out = sim;
	return NBODY_OK;
}

}  // extern "C"
