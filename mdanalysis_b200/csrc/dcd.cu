// dcd.cu -- bulk reader for CHARMM / NAMD DCD trajectories (host code only).
//
// The step in front of the hot path: the reference reads a DCD
// one frame at a time through DCDReader._read_next_timestep (coordinates/DCD.py:197-209) or in
// bulk through DCDFile.readframes (lib/formats/libdcd.pyx:583-700, C side lib/formats/include/
// readdcd.h).  This is the bulk path, written for the GPU pipeline: frames [start, stop, step)
// of the selected atoms land in ONE caller-owned float32 [n, nsel, 3] buffer ('fac' order, what
// the kernels consume; the caller passes pinned memory so the H2D copy that follows runs at PCIe
// rate), the unit cells in a double [n, 6] buffer in the file's own order
// [A, gamma, B, beta, alpha, C] -- the conversion to [A, B, C, alpha, beta, gamma] is the
// caller's, as in the reference (DCD.py:225-256).
//
// File layout (Fortran unformatted records, 4-byte markers, either endianness):
//   84 | 'CORD' icntrl[20] | 84     icntrl: 0 NSET, 1 ISTART, 2 NSAVC, 8 NAMNF (fixed atoms),
//                                           9 DELTA (float32), 10 unit cell?, 11 4th dimension?, 19 CHARMM version
//   len | ntitle, 80-char lines | len
//   4 | natoms | 4
//   per frame: [48 | 6 doubles | 48]  4N | x | 4N  4N | y | 4N  4N | z | 4N  [4N | w | 4N]
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <vector>

#include "common.cuh"

struct mdagpu_dcd {
    int fd = -1;
    bool swap = false;
    int64_t natoms = 0, nframes = 0;
    int has_cell = 0, has_4d = 0, charmm = 0;
    int32_t istart = 0, nsavc = 0, nset = 0, namnf = 0;
    double delta = 0.0;
    int64_t first = 0, frame_bytes = 0;        // offset of frame 0, bytes per frame
    std::vector<unsigned char> buf;            // one frame as stored
};

using namespace mdagpu;

static inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
static inline uint64_t bswap64(uint64_t v) { return __builtin_bswap64(v); }

static bool read_at(int fd, void *dst, size_t n, int64_t off)
{
    unsigned char *p = (unsigned char *)dst;
    while (n) {
        const ssize_t r = pread(fd, p, n, off);
        if (r <= 0) {
            if (r < 0 && errno == EINTR) continue;
            return false;
        }
        p += r; n -= (size_t)r; off += r;
    }
    return true;
}

extern "C" int mdagpu_dcd_open(const char *path, mdagpu_dcd **out)
{
    MDA_REQUIRE(path && out, "mdagpu_dcd_open: NULL argument");
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return set_error(MDAGPU_ERR_INVALID, "DCD: cannot open %s: %s", path, strerror(errno));
    mdagpu_dcd *d = new mdagpu_dcd;
    d->fd = fd;
    auto fail = [&](const char *why) {
        close(fd);
        delete d;
        return set_error(MDAGPU_ERR_INVALID, "DCD: %s: %s", path, why);
    };
    uint32_t head[23];
    if (!read_at(fd, head, sizeof head, 0)) return fail("file too short for a header");
    if (head[0] == 84u) d->swap = false;
    else if (bswap32(head[0]) == 84u) d->swap = true;
    else return fail("not a DCD file (first record is not 84 bytes long; 64-bit record markers are not supported)");
    if (memcmp(&head[1], "CORD", 4) != 0) return fail("not a coordinate DCD file (no CORD magic)");
    auto i32 = [&](uint32_t v) { return (int32_t)(d->swap ? bswap32(v) : v); };
    const uint32_t *ic = head + 2;
    d->nset = i32(ic[0]); d->istart = i32(ic[1]); d->nsavc = i32(ic[2]); d->namnf = i32(ic[8]);
    d->charmm = i32(ic[19]);
    if (d->charmm) {
        uint32_t raw = d->swap ? bswap32(ic[9]) : ic[9];
        float fl;
        memcpy(&fl, &raw, 4);
        d->delta = (double)fl;
        d->has_cell = i32(ic[10]) != 0;
        d->has_4d = i32(ic[11]) != 0;
    } else {                                    // X-PLOR: DELTA is a double in icntrl[9..10]
        uint64_t raw;
        memcpy(&raw, &ic[9], 8);
        if (d->swap) raw = bswap64(raw);
        memcpy(&d->delta, &raw, 8);
    }
    if (i32(head[22]) != 84) return fail("corrupt header (closing record marker)");
    int64_t off = 92;
    uint32_t len;
    if (!read_at(fd, &len, 4, off)) return fail("truncated title block");
    const int64_t tlen = i32(len);
    if (tlen < 4 || (tlen - 4) % 80 != 0) return fail("corrupt title block");
    off += 4 + tlen + 4;
    uint32_t nat[3];
    if (!read_at(fd, nat, 12, off)) return fail("truncated atom-count block");
    if (i32(nat[0]) != 4 || i32(nat[2]) != 4) return fail("corrupt atom-count block");
    d->natoms = i32(nat[1]);
    off += 12;
    if (d->natoms <= 0) return fail("no atoms");
    if (d->namnf != 0) return fail("trajectories with fixed atoms (NAMNF > 0) are not supported");
    d->first = off;
    const int64_t plane = 8 + 4 * d->natoms;
    d->frame_bytes = (d->has_cell ? 56 : 0) + (3 + (d->has_4d ? 1 : 0)) * plane;
    struct stat st;
    if (fstat(fd, &st) != 0) return fail("fstat failed");
    // like the reference, trust the file size rather than NSET (libdcd.pyx: _estimate_n_frames)
    d->nframes = (st.st_size - d->first) / d->frame_bytes;
    d->buf.resize((size_t)d->frame_bytes);
    *out = d;
    return MDAGPU_OK;
}

extern "C" void mdagpu_dcd_close(mdagpu_dcd *d)
{
    if (!d) return;
    if (d->fd >= 0) close(d->fd);
    delete d;
}

extern "C" int mdagpu_dcd_header(const mdagpu_dcd *d, int64_t *natoms, int64_t *nframes, int *has_cell,
                                 int32_t *istart, int32_t *nsavc, double *delta, int *charmm)
{
    MDA_REQUIRE(d, "mdagpu_dcd_header: NULL handle");
    if (natoms) *natoms = d->natoms;
    if (nframes) *nframes = d->nframes;
    if (has_cell) *has_cell = d->has_cell;
    if (istart) *istart = d->istart;
    if (nsavc) *nsavc = d->nsavc;
    if (delta) *delta = d->delta;
    if (charmm) *charmm = d->charmm;
    return MDAGPU_OK;
}

extern "C" int mdagpu_dcd_readframes(mdagpu_dcd *d, int64_t start, int64_t stop, int64_t step,
                                     const int64_t *indices, int64_t nsel, float *xyz, double *unitcell)
{
    MDA_REQUIRE(d && xyz, "mdagpu_dcd_readframes: NULL argument");
    MDA_REQUIRE(step != 0, "mdagpu_dcd_readframes: step is zero");
    const int64_t N = d->natoms;
    if (!indices) nsel = N;
    for (int64_t k = 0; indices && k < nsel; ++k)
        MDA_REQUIRE(indices[k] >= 0 && indices[k] < N, "mdagpu_dcd_readframes: atom index %lld out of range",
                    (long long)indices[k]);
    const int64_t plane = 8 + 4 * N;
    int64_t n = 0;
    for (int64_t fr = start; step > 0 ? fr < stop : fr > stop; fr += step, ++n) {
        MDA_REQUIRE(fr >= 0 && fr < d->nframes, "mdagpu_dcd_readframes: frame %lld outside [0, %lld)",
                    (long long)fr, (long long)d->nframes);
        unsigned char *b = d->buf.data();
        if (!read_at(d->fd, b, (size_t)d->frame_bytes, d->first + fr * d->frame_bytes))
            return set_error(MDAGPU_ERR_INVALID, "DCD: short read at frame %lld", (long long)fr);
        if (d->has_cell) {
            uint32_t m;
            memcpy(&m, b, 4);
            if ((int32_t)(d->swap ? bswap32(m) : m) != 48)
                return set_error(MDAGPU_ERR_INVALID, "DCD: corrupt unit-cell record at frame %lld", (long long)fr);
            if (unitcell)
                for (int k = 0; k < 6; ++k) {
                    uint64_t raw;
                    memcpy(&raw, b + 4 + 8 * k, 8);
                    if (d->swap) raw = bswap64(raw);
                    memcpy(&unitcell[6 * n + k], &raw, 8);
                }
            b += 56;
        } else if (unitcell) {
            for (int k = 0; k < 6; ++k) unitcell[6 * n + k] = 0.0;
        }
        uint32_t m;
        memcpy(&m, b, 4);
        if ((int64_t)(int32_t)(d->swap ? bswap32(m) : m) != 4 * N)
            return set_error(MDAGPU_ERR_INVALID, "DCD: corrupt coordinate record at frame %lld", (long long)fr);
        const uint32_t *px = (const uint32_t *)(b + 4), *py = (const uint32_t *)(b + plane + 4),
                       *pz = (const uint32_t *)(b + 2 * plane + 4);
        uint32_t *dst = (uint32_t *)(xyz + (size_t)n * (size_t)nsel * 3);
        if (!indices && !d->swap) {
            for (int64_t i = 0; i < N; ++i) { dst[3 * i] = px[i]; dst[3 * i + 1] = py[i]; dst[3 * i + 2] = pz[i]; }
        } else {
            for (int64_t k = 0; k < nsel; ++k) {
                const int64_t i = indices ? indices[k] : k;
                uint32_t x = px[i], y = py[i], z = pz[i];
                if (d->swap) { x = bswap32(x); y = bswap32(y); z = bswap32(z); }
                dst[3 * k] = x; dst[3 * k + 1] = y; dst[3 * k + 2] = z;
            }
        }
    }
    return MDAGPU_OK;
}
