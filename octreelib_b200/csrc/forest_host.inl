// Host-side orchestration of the Forest (included at the end of forest.cu).
namespace ol {

static inline unsigned nblk(size_t n, int t = 256) { return n ? (unsigned)((n + t - 1) / t) : 1u; }

// Page-locked scratch blocks (256 B) for scalar read-backs.  cudaMallocHost / cudaFreeHost cost 10-200 ms each once
// gigabytes of device memory are mapped (measured: 24 ms of a 31 ms step went into creating the forest), so the blocks
// are recycled process-wide instead of being allocated per forest; they are never returned to the driver.
static std::mutex g_pinned_mutex;
static std::vector<void*> g_pinned_free;
static void* pinned_scratch_get() {
    {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        if (!g_pinned_free.empty()) {
            void* p = g_pinned_free.back();
            g_pinned_free.pop_back();
            return p;
        }
    }
    void* p = nullptr;
    OL_CUDA(cudaMallocHost(&p, 256));
    return p;
}
static void pinned_scratch_put(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    g_pinned_free.push_back(p);
}

// upload streams of the pinned-source inserts (process-wide, created on first use) and the events that order them against
// a forest's own stream
constexpr int UPLOAD_MAX_DEVICES = 64;
struct UploadLane {
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t events[3] = {nullptr, nullptr, nullptr};
};
static UploadLane g_upload[UPLOAD_MAX_DEVICES];
static UploadLane& upload_lane(int device) {
    OL_REQUIRE(device >= 0 && device < UPLOAD_MAX_DEVICES, OL_ERR_INVALID, "device index out of range");
    UploadLane& u = g_upload[device];
    if (!u.streams[0]) {
        for (auto& s : u.streams) OL_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        for (auto& e : u.events) OL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return u;
}

// next upload stream; the first copy after work on the forest's stream that the copies must follow (allocation, growth)
// makes both upload streams wait for that work
cudaStream_t Forest::upload_stream() {
    UploadLane& u = upload_lane(cfg.device);
    if (!uploads_pending) {
        OL_CUDA(cudaEventRecord(u.events[2], ctx.stream));
        for (auto s : u.streams) OL_CUDA(cudaStreamWaitEvent(s, u.events[2], 0));
        uploads_pending = true;
    }
    upload_turn ^= 1;
    return u.streams[upload_turn];
}

// the forest's stream waits for every copy issued on the upload streams
void Forest::join_uploads() {
    if (!uploads_pending) return;
    UploadLane& u = upload_lane(cfg.device);
    for (int k = 0; k < 2; ++k) {
        OL_CUDA(cudaEventRecord(u.events[k], u.streams[k]));
        OL_CUDA(cudaStreamWaitEvent(ctx.stream, u.events[k], 0));
    }
    uploads_pending = false;
}

static size_t g_last_build_points = 0;  // points of the last grid built by this process (capacity guess of the next one)

Forest::Forest(const ol_forest_config& c) : cfg(c) {
    OL_REQUIRE(c.voxel_edge_length > 0 && std::isfinite(c.voxel_edge_length), OL_ERR_INVALID,
               "voxel_edge_length must be a positive finite number");
    max_depth = c.max_depth <= 0 ? OL_MAX_DEPTH : c.max_depth;
    OL_REQUIRE(max_depth <= OL_MAX_DEPTH, OL_ERR_INVALID, "max_depth must be <= 21");
    OL_CUDA(cudaSetDevice(c.device));
    ctx.stream = (cudaStream_t)c.stream;
    ctx.alloc_fn = c.alloc;
    ctx.free_fn = c.free;
    ctx.alloc_user = c.alloc_user;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
    ctx.num_sms = sms;
    d_bbox.reset(ctx, 6);
    d_err.reset(ctx, 1);
    d_err.zero();
    ctx.d_err = d_err.get();
    ctx.prof = &prof;
    long long init[6] = {LLONG_MAX, LLONG_MAX, LLONG_MAX, LLONG_MIN, LLONG_MIN, LLONG_MIN};
    OL_CUDA(cudaMemcpyAsync(d_bbox.get(), init, sizeof(init), cudaMemcpyHostToDevice, ctx.stream));
    pinned = pinned_scratch_get();
    mailbox = pinned_scratch_get();
    // no synchronisation: the initial values are pageable sources (staged when cudaMemcpyAsync returns)
    seg_start.push_back(0);
}

Forest::~Forest() {
    if (uploads_pending)
        for (auto s : g_upload[cfg.device].streams) cudaStreamSynchronize(s);  // copies into the point array that is about to be released
    cudaStreamSynchronize(ctx.stream);  // also: no kernel may post into the mailbox after it went back to the pool
    pinned_scratch_put(pinned);
    pinned_scratch_put(mailbox);
}

Mail Forest::mail_open(int slot, const uint32_t* aux) {
    Mail m;
    m.slot = reinterpret_cast<volatile unsigned long long*>(mailbox) + 4 * (size_t)slot;
    m.ticket = ++g_mail_ticket;
    m.err = d_err.get();
    m.aux = aux;
    return m;
}

void Forest::resolve_cells() {
    if (!cells_pending) return;
    cells_pending = false;
    const MailResult r = mail_take(cells_mail);
    try {
        throw_device_errors(r.err);  // keygen range checks
    } catch (...) {
        built = false;
        throw;
    }
    C = (uint32_t)r.total;
}

void Forest::resolve_blocks() {
    if (!blocks_pending) return;
    blocks_pending = false;
    NB = (uint32_t)mail_take(blocks_mail).total;
}

uint32_t Forest::read_u32(const uint32_t* dptr) {
    uint32_t v = 0;
    read_back({{dptr, 4, &v}});
    return v;
}
unsigned long long Forest::read_u64(const unsigned long long* dptr) {
    unsigned long long v = 0;
    read_back({{dptr, 8, &v}});
    return v;
}

// several scalars, one synchronisation (every stream synchronisation costs the GPU ~20 us of idle time: the host has to
// wake up, read, and issue the next kernels)
void Forest::read_back(std::initializer_list<ReadItem> items) {
    size_t off = 0;
    for (const auto& it : items) {
        OL_REQUIRE(off + it.bytes <= 256, OL_ERR_INTERNAL, "read_back: scratch block too small");
        OL_CUDA(cudaMemcpyAsync(static_cast<char*>(pinned) + off, it.src, it.bytes, cudaMemcpyDeviceToHost, ctx.stream));
        off += (it.bytes + 7) & ~(size_t)7;
    }
    ctx.sync();
    off = 0;
    for (const auto& it : items) {
        memcpy(it.dst, static_cast<char*>(pinned) + off, it.bytes);
        off += (it.bytes + 7) & ~(size_t)7;
    }
}

// device error word -> exception (the word is cleared first); the RANSAC bits are not errors and stay
void Forest::throw_device_errors(uint32_t e) {
    const uint32_t fatal = e & (DEVERR_NONFINITE | DEVERR_CELL_RANGE | DEVERR_OUT_OF_NODE | DEVERR_DEPTH_CAP);
    if (!fatal) return;
    const uint32_t rest = e & ~fatal;
    OL_CUDA(cudaMemcpyAsync(d_err.get(), &rest, 4, cudaMemcpyHostToDevice, ctx.stream));
    ctx.sync();
    if (e & DEVERR_NONFINITE) throw Error{OL_ERR_NONFINITE, "point cloud contains NaN or infinite coordinates"};
    if (e & DEVERR_CELL_RANGE) throw Error{OL_ERR_RANGE, "cell coordinates out of the representable range"};
    if (e & DEVERR_OUT_OF_NODE)
        throw Error{OL_ERR_OUT_OF_NODE,
                    "a point lies outside the octree node that is being subdivided (the reference raises IndexError "
                    "or mis-routes it, octree/octree.py:94-98)"};
    if (e & DEVERR_DEPTH_CAP)
        throw Error{OL_ERR_DEPTH_CAP, "subdivision criterion still true at the maximum octree depth " +
                                          std::to_string(max_depth) + " (the reference would keep recursing)"};
}

void Forest::check_device_errors() { throw_device_errors(read_u32(d_err.get())); }

void Forest::upload_segments() {
    // ONE allocation and ONE copy for the three tables: [seg_first (int64) | seg_start (uint32, + sentinel) | seg_pose (int32)];
    // d_seg_* are views into it (DevBuf without a context: nothing to free)
    const size_t S = seg_pose.size(), Sp = S ? S : 1;
    const size_t off_start = Sp * 8, off_pose = off_start + (S + 1) * 4, bytes = off_pose + Sp * 4;
    std::vector<unsigned char> blob(bytes, 0);
    if (S) {
        memcpy(blob.data(), seg_first.data(), S * 8);
        memcpy(blob.data() + off_pose, seg_pose.data(), S * 4);
    }
    memcpy(blob.data() + off_start, seg_start.data(), (S + 1) * 4);
    seg_blob.reset(ctx, bytes);
    h2d(ctx, seg_blob.get(), blob.data(), bytes);
    // no synchronisation: cudaMemcpyAsync from pageable memory returns once the source has been staged
    auto view = [](auto& buf, void* p, size_t count) {
        buf.release();
        buf.ctx = nullptr;
        buf.ptr = static_cast<decltype(buf.ptr)>(p);
        buf.count = count;
    };
    view(d_seg_first, seg_blob.get(), Sp);
    view(d_seg_start, seg_blob.get() + off_start, S + 1);
    view(d_seg_pose, seg_blob.get() + off_pose, Sp);
}

// ---------------------------------------------------------------------------------------------
// Grid.insert_points (grid.py:58-109): upload + bounding box only; the grouping is deferred to
// build() so that all poses are keyed and sorted in one pass.
// ---------------------------------------------------------------------------------------------
int Forest::insert(const double* xyz, int64_t n, bool on_device, const int64_t* seg_sizes, const int32_t* seg_pose_in,
                   const int64_t* seg_first_in, int n_segments, int n_poses_total) {
    OL_REQUIRE(n >= 0, OL_ERR_INVALID, "negative point count");
    OL_REQUIRE(N + (size_t)n < (1ull << 31), OL_ERR_INVALID, "more than 2^31 - 1 points per forest are not supported");
    materialize_snapshot();
    ensure_alive();
    if (shaped && I > 0) save_shape();  // a later pose follows the existing subdivision (octree_manager.py:171)
    if (q_known) {  // points after an adopted buffer: the bounding box is taken over everything again
        q_known = false;
        bbox_done = 0;
    }
    if (N + (size_t)n > cap) {
        join_uploads();  // copies still in flight on the upload streams target the old array
        size_t ncap = std::max<size_t>(N + (size_t)n, cap + cap / 2 + 1024);
        // first guess: the size of the previous map of this process (see insert_batch) - once this map has reached 1/32 of it,
        // so that a small grid after a large one does not reserve the large one's memory (the geometric growth up to that
        // point copies a few percent of the final size)
        if (g_last_build_points >= N + (size_t)n && N + (size_t)n >= g_last_build_points / 32) ncap = std::max(ncap, g_last_build_points);
        DevBuf<double> np(ctx, ncap * 3);
        d2d(ctx, np.get(), P64.get(), N * 3);
        P64.swap(np);
        if (alive_r.get()) {  // only exists once something was removed (apply_keep)
            DevBuf<uint8_t> na(ctx, ncap);
            d2d(ctx, na.get(), alive_r.get(), N);
            OL_CUDA(cudaMemsetAsync(na.get() + N, 1, ncap - N, ctx.stream));
            alive_r.swap(na);
        }
        cap = ncap;
    }
    // one copy per insert; the bounding box / non-finite check of the new points runs once, in build()
    bool src_pinned = false;
    if (!on_device && n > 0) {
        cudaPointerAttributes attr{};
        src_pinned = cudaPointerGetAttributes(&attr, xyz) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!src_pinned) cudaGetLastError();  // an unregistered pointer may leave a sticky-free error code behind on old drivers
    }
    if (n > 0) {
        if (src_pinned) {
            // Page-locked sources go through TWO upload streams in turn: the DMA set-up of one copy overlaps the transfer
            // of the other (838 copies of 2.9 MB on one stream: 4 us bubble each, 51 GB/s while busy).  build() joins them.
            cudaStream_t up = upload_stream();
            OL_CUDA(cudaMemcpyAsync(P64.get() + N * 3, xyz, (size_t)n * 24, cudaMemcpyHostToDevice, up));
        } else {
            OL_CUDA(cudaMemcpyAsync(P64.get() + N * 3, xyz, (size_t)n * 24, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                    ctx.stream));
        }
    }
    int pose_index;
    if (n_segments <= 0) {
        pose_index = n_poses;
        seg_pose.push_back(pose_index);
        seg_first.push_back(0);
        seg_start.back() = (uint32_t)N;
        seg_start.push_back((uint32_t)(N + n));
        n_poses += 1;
    } else {
        pose_index = -1;
        size_t r = N;
        int64_t tot = 0;
        for (int s = 0; s < n_segments; ++s) {
            OL_REQUIRE(seg_sizes[s] >= 0 && seg_pose_in[s] >= 0 && seg_pose_in[s] < n_poses_total, OL_ERR_INVALID,
                       "bad segment description");
            if (!seg_pose.empty() && seg_pose_in[s] < seg_pose.back()) segs_pose_monotone = false;
            seg_pose.push_back(seg_pose_in[s]);
            seg_first.push_back(seg_first_in ? seg_first_in[s] : 0);
            seg_start.back() = (uint32_t)r;
            r += (size_t)seg_sizes[s];
            seg_start.push_back((uint32_t)r);
            tot += seg_sizes[s];
        }
        OL_REQUIRE(tot == n, OL_ERR_INVALID, "segment sizes do not add up to n");
        n_poses = std::max(n_poses, n_poses_total);
    }
    N += (size_t)n;
    if ((int)pose_epoch.size() < n_poses) pose_epoch.resize(n_poses, n_subdivide_calls);
    built = false;
    shaped = false;
    order_valid = blocks_valid = ransac_valid = false;
    if (!on_device && n > 0 && !src_pinned) {
        // Pageable host memory: cudaMemcpyAsync returns once the source has been staged, so the caller may reuse it.
        // Page-locked (pinned) host memory is read by DMA later: the caller deliberately handed over an asynchronous
        // buffer and must leave it unchanged until the next call that returns results (documented in Grid.insert_points);
        // waiting here would serialise 839 copies of a 100 M-point map with ~25 us bubbles each.
        ctx.sync();
    }
    return pose_index;
}

// Multi-GPU exchange: the receive buffer becomes the point array of an empty forest (no copy).  The (source rank, pose)
// runs arrive as segments; q_lo / q_hi bound the cell coordinates of everything in the buffer, so build() needs neither a
// bounding-box pass nor its read-back.  The buffer must stay valid until the forest is destroyed or disown_points() ran.
void Forest::adopt_points(double* ext, size_t n, const int64_t* seg_sizes, const int32_t* seg_pose_in, const int64_t* seg_first_in,
                          int n_segments, int n_poses_total, const long long qlo[3], const long long qhi[3]) {
    OL_REQUIRE(N == 0 && n_poses == 0 && !shaped, OL_ERR_STATE, "only an empty forest can adopt a point array");
    OL_REQUIRE(n < (1ull << 31), OL_ERR_INVALID, "more than 2^31 - 1 points per forest are not supported");
    P64.release();
    P64.ctx = nullptr;
    P64.ptr = ext;
    P64.count = n * 3;
    cap = n;
    size_t r = 0;
    for (int s = 0; s < n_segments; ++s) {
        OL_REQUIRE(seg_sizes[s] >= 0 && seg_pose_in[s] >= 0 && seg_pose_in[s] < n_poses_total, OL_ERR_INVALID, "bad segment description");
        if (!seg_pose.empty() && seg_pose_in[s] < seg_pose.back()) segs_pose_monotone = false;
        seg_pose.push_back(seg_pose_in[s]);
        seg_first.push_back(seg_first_in ? seg_first_in[s] : 0);
        seg_start.back() = (uint32_t)r;
        r += (size_t)seg_sizes[s];
        seg_start.push_back((uint32_t)r);
    }
    OL_REQUIRE(r == n, OL_ERR_INVALID, "segment sizes do not add up to n");
    n_poses = n_poses_total;
    N = n;
    bbox_done = n;
    q_known = true;
    for (int a = 0; a < 3; ++a) {
        q_lo[a] = qlo[a];
        q_hi[a] = qhi[a];
    }
    if ((int)pose_epoch.size() < n_poses) pose_epoch.resize(n_poses, n_subdivide_calls);
    built = false;
    shaped = false;
    order_valid = blocks_valid = ransac_valid = false;
}

void Forest::disown_points() {
    if (!points_external()) return;
    DevBuf<double> own(ctx, std::max<size_t>(cap, 1) * 3);
    d2d(ctx, own.get(), P64.get(), N * 3);
    P64.swap(own);  // `own` now holds the foreign pointer without a context: its destructor frees nothing
    ctx.sync();     // the caller may hand the buffer to somebody else as soon as this returns
}

// Many device-resident poses at once (the Python host defers `insert_points` of CUDA tensors and flushes them here):
// one exact-size growth of the point array, one pointer-table upload, one copy kernel - instead of a driver call and a
// possible re-allocation per pose.  Every cloud becomes one new pose; returns the index of the first one.
int Forest::insert_batch(const double* const* xyz_dev, const int64_t* sizes, int count) {
    OL_REQUIRE(count >= 0, OL_ERR_INVALID, "negative batch size");
    const int first_pose = n_poses;
    if (count == 0) return first_pose;
    size_t total = 0;
    for (int c = 0; c < count; ++c) {
        OL_REQUIRE(sizes[c] >= 0, OL_ERR_INVALID, "negative point count");
        total += (size_t)sizes[c];
    }
    OL_REQUIRE(N + total < (1ull << 31), OL_ERR_INVALID, "more than 2^31 - 1 points per forest are not supported");
    materialize_snapshot();
    ensure_alive();
    if (shaped && I > 0) save_shape();
    if (q_known) {
        q_known = false;
        bbox_done = 0;
    }
    join_uploads();  // (mixed use: host arrays inserted before this batch)
    if (N + total > cap) {
        // The host hands its clouds over in several batches while it is still inserting (forest.py flushes every 128 poses,
        // so that the copy runs while the host language works through the remaining insert calls).  The final size is
        // not known yet: the point count of the previous build of this process is the first guess (a pipeline builds
        // maps of similar size step after step: no growth copy and no slack in the steady state), geometric growth
        // otherwise.
        size_t ncap = N + total;
        const size_t guess = g_last_build_points;
        if (guess >= ncap && ncap >= guess / 32)
            ncap = guess;
        else if (cap > 0)
            ncap = std::max(ncap, cap * 2);
        DevBuf<double> np(ctx, ncap * 3);
        d2d(ctx, np.get(), P64.get(), N * 3);
        P64.swap(np);
        if (alive_r.get()) {
            DevBuf<uint8_t> na(ctx, ncap);
            d2d(ctx, na.get(), alive_r.get(), N);
            OL_CUDA(cudaMemsetAsync(na.get() + N, 1, ncap - N, ctx.stream));
            alive_r.swap(na);
        }
        cap = ncap;
    }
    // chunk length: at least ~16 chunks per SM (a batch of 128 poses is small next to the whole map: with full-size chunks it
    // was 5 CTAs per SM and ran at 60 % of the bandwidth of the single big launch), a multiple of 3, at most INSERT_CHUNK
    unsigned long long chunk_len = (unsigned long long)total * 3ull / ((unsigned long long)ctx.num_sms * 16ull);
    chunk_len = std::min<unsigned long long>(std::max<unsigned long long>(chunk_len / 3ull * 3ull, 3ull * 1024ull), INSERT_CHUNK);
    std::vector<InsertCloud> rows;
    rows.reserve((size_t)count);
    size_t r = N;
    unsigned long long dst = 0, n_chunks = 0;
    for (int c = 0; c < count; ++c) {
        const unsigned long long len = (unsigned long long)sizes[c] * 3ull;
        if (len) rows.push_back(InsertCloud{xyz_dev[c], dst, len, (unsigned)n_chunks, 0u});
        n_chunks += (len + chunk_len - 1) / chunk_len;
        dst += len;
        seg_pose.push_back(n_poses);
        seg_first.push_back(0);
        seg_start.back() = (uint32_t)r;
        r += (size_t)sizes[c];
        seg_start.push_back((uint32_t)r);
        n_poses += 1;
    }
    if (n_chunks) {
        DevBuf<InsertCloud> d_rows(ctx, rows.size());
        h2d(ctx, d_rows.get(), rows.data(), rows.size());
        ProfScope ps(ctx, "insert_batch", (double)total);
        insert_batch_kernel<<<(unsigned)n_chunks, INSERT_THREADS, 0, ctx.stream>>>(d_rows.get(), (int)rows.size(), (unsigned)chunk_len,
                                                                                   P64.get() + N * 3, d_bbox.get(), d_err.get());
        OL_CHECK_LAUNCH();
        // No synchronisation: the pageable cloud table has been staged when cudaMemcpyAsync returns, and the caller's
        // clouds are torch tensors whose memory is recycled in stream order on this same stream (documented in
        // include/octreelib_b200.h: sources must stay valid until the work enqueued here has run).
    }
    if (bbox_done == N) bbox_done = N + total;  // the copy kernel already folded the batch into the bounding box
    N += total;
    if ((int)pose_epoch.size() < n_poses) pose_epoch.resize(n_poses, n_subdivide_calls);
    built = false;
    shaped = false;
    order_valid = blocks_valid = ransac_valid = false;
    return first_pose;
}

// ---------------------------------------------------------------------------------------------
// K1-K3: keys, sort, cells, (cell, pose) table
// ---------------------------------------------------------------------------------------------
void Forest::build() {
    if (!built) build_enqueue();
    resolve_cells();
}

void Forest::build_enqueue() {
    if (built) return;
    join_uploads();
    cells_pending = false;
    g_last_build_points = N;
    if (bbox_done < N) {  // K0 over everything inserted since the last build
        const size_t m = N - bbox_done;
        unsigned g = std::min<unsigned>(nblk(m * 3, BBOX_THREADS), (unsigned)ctx.num_sms * 10);
        ProfScope ps(ctx, "bbox", (double)m);
        bbox_kernel<<<g, BBOX_THREADS, 0, ctx.stream>>>(P64.get() + bbox_done * 3, m, d_bbox.get(), d_err.get());
        OL_CHECK_LAUNCH();
        bbox_done = N;
    }
    upload_segments();
    const int S = (int)seg_pose.size();
    kp = KeyParams{};
    kp.edge = cfg.voxel_edge_length;
    kp.inv_edge = pow2_reciprocal(kp.edge);
    for (int a = 0; a < 3; ++a) kp.corner[a] = cfg.corner[a];
    kp.single_cell = cfg.single_cell;
    kp.depth = std::min(max_depth, MORTON_INITIAL_DEPTH);  // deeper levels are computed on demand (extend_morton)
    mort32 = kp.depth <= MORTON32_MAX_DEPTH;
    kp.pose_bits = segs_pose_monotone ? 0 : bit_length_u64((uint64_t)std::max(n_poses - 1, 0));
    int bits[3] = {0, 0, 0};
    // The point-sized work buffers are allocated BEFORE the read-back below, while the GPU is still busy with the insert
    // copy: every large allocation is a trip into the host allocator callback (5-10 us), and nine of them between the
    // synchronisation and the key kernel left the GPU idle for 50-65 us.  The key width is only known after the
    // read-back; the buffers are sized by the width of the previous build of this process and re-made if it differs.
    const uint32_t n_pre = (uint32_t)N;
    static int s_last_key_bytes = 4;
    const int pre_key_bytes = s_last_key_bytes;
    DevBuf<uint64_t> kraw0, kraw1, mort_r;
    DevBuf<uint32_t> vals0, vals1;
    DevBuf<unsigned long long> d_total;
    if (n_pre) {
        cellidx0.reset(ctx, n_pre);
        kraw0.reset(ctx, pre_key_bytes == 4 ? ((size_t)n_pre + 1) / 2 : (size_t)n_pre);
        kraw1.reset(ctx, pre_key_bytes == 4 ? ((size_t)n_pre + 1) / 2 : (size_t)n_pre);
        mort_r.reset(ctx, mort_len(n_pre));
        vals0.reset(ctx, n_pre);
        vals1.reset(ctx, n_pre);
        d_total.reset(ctx, 1);
    }
    if (q_known && bbox_done == N) {  // adopted receive buffer: the cell-coordinate range came with it, nothing to read back
        if (!cfg.single_cell && N > 0)
            for (int a = 0; a < 3; ++a) {
                kp.qmin[a] = q_lo[a];
                bits[a] = bit_length_u64((uint64_t)(q_hi[a] - q_lo[a]));
            }
    } else {  // the build's one wait for the device: bounding box + error word, posted into the mailbox (slots 6-7)
        long long hb[6];
        uint32_t e = 0;
        {
            Mail m = mail_open(6);
            post_bbox_kernel<<<1, 32, 0, ctx.stream>>>(d_bbox.get(), m);
            OL_CHECK_LAUNCH();
            mail_take(m);
            for (int a = 0; a < 6; ++a) hb[a] = (long long)m.slot[1 + a];
            e = (uint32_t)m.slot[7];
        }
        throw_device_errors(e);
        if (!cfg.single_cell && N > 0) {
            for (int a = 0; a < 3; ++a) {
                double lo = cell_coord(ordered_to_double(hb[a]), kp.corner[a], kp.edge);
                double hi = cell_coord(ordered_to_double(hb[3 + a]), kp.corner[a], kp.edge);
                OL_REQUIRE(std::fabs(lo) < 4503599627370496.0 && std::fabs(hi) < 4503599627370496.0, OL_ERR_RANGE,
                           "cell coordinates exceed 2^52");
                kp.qmin[a] = (long long)lo;
                bits[a] = bit_length_u64((uint64_t)((long long)hi - (long long)lo));
            }
        }
    }
    const int cell_bits = bits[0] + bits[1] + bits[2];
    key_bits = cell_bits + kp.pose_bits;
    // Morton levels that fit into a 32-bit sort key below the packed cell key (+ 1 bit for the out-of-node flag): they ride
    // through the sort for free.  Deeper levels are computed on demand (extend_morton), like the levels beyond
    // MORTON_INITIAL_DEPTH of the separate-array layout.
    static const bool no_embed = getenv("OL_NO_EMBED") != nullptr;
    key_embed = 0;
    if (!no_embed && N > 0 && key_bits + 1 + 3 * std::min(kp.depth, 3) <= 32) {
        key_embed = std::min(kp.depth, (32 - key_bits - 1) / 3);
        kp.depth = key_embed;
    }
    const int fw = key_embed ? 3 * key_embed + 1 : 0;
    OL_REQUIRE(key_bits <= 64, OL_ERR_RANGE,
               "the grid spans too many cells: packed cell key needs " + std::to_string(key_bits) + " bits (max 64)");
    kp.shift[2] = kp.pose_bits;
    kp.shift[1] = kp.shift[2] + bits[2];
    kp.shift[0] = kp.shift[1] + bits[1];

    const uint32_t n = (uint32_t)N;
    C = 0;
    CP = 0;
    A0 = n;
    if (n == 0) {
        perm0.reset(ctx, 0);
        mort0.reset(ctx, 0);
        cellidx0.reset(ctx, 0);
        cell_key.reset(ctx, 0);
        cell_start0.reset(ctx, 1);
        cell_start0.zero();
        cp_cell.reset(ctx, 0);
        cp_pose.reset(ctx, 0);
        cell_first_pose.reset(ctx, 0);
        cp_valid = true;
        built = true;
        base_dirty = false;
        return;
    }
    // the cell tables are sized by what the key space allows (the number of cells is only known after K3)
    const size_t c_max = cell_bits < 31 ? std::min<size_t>(n, (size_t)1 << cell_bits) : n;
    cell_key.reset(ctx, c_max);
    cell_start0.reset(ctx, c_max + 1);
    const int key_bytes = key_bits + fw <= 32 ? 4 : 8;
    if (key_bytes != pre_key_bytes) {
        kraw0.reset(ctx, key_bytes == 4 ? ((size_t)n + 1) / 2 : (size_t)n);
        kraw1.reset(ctx, key_bytes == 4 ? ((size_t)n + 1) / 2 : (size_t)n);
    }
    s_last_key_bytes = key_bytes;
    // K1 + K2 + K3 for one key width
    auto run = [&](auto key_tag) {
        using KeyT = decltype(key_tag);
        struct Raw {
            DevBuf<uint64_t>& b;
            KeyT* get() const { return reinterpret_cast<KeyT*>(b.get()); }
            void swap(Raw& o) { b.swap(o.b); }
            void release() { b.release(); }
        } keys0{kraw0}, keys1{kraw1};
        {
            ProfScope ps(ctx, "keygen", (double)n);
            if (mort32)
                keygen_kernel<KeyT, uint32_t><<<nblk(n), 256, 0, ctx.stream>>>(P64.get(), n, kp, d_seg_start.get(), d_seg_pose.get(), S,
                                                                               keys0.get(), vals0.get(),
                                                                               reinterpret_cast<uint32_t*>(mort_r.get()), d_err.get(),
                                                                               key_embed);
            else
                keygen_kernel<KeyT, uint64_t><<<nblk(n), 256, 0, ctx.stream>>>(P64.get(), n, kp, d_seg_start.get(), d_seg_pose.get(), S,
                                                                               keys0.get(), vals0.get(), mort_r.get(), d_err.get(), 0);
            OL_CHECK_LAUNCH();
        }
        const int which = radix_sort_pairs<KeyT>(ctx, keys0.get(), keys1.get(), vals0.get(), vals1.get(), n, fw, fw + key_bits, true);
        if (which) {
            keys0.swap(keys1);
            vals0.swap(vals1);
        }
        keys1.release();
        vals1.release();
        // K3: cells = runs of equal cell key
        // K3: the number of cells is POSTED to the host by the kernel (the host reads it when it needs it: resolve_cells),
        // and the kernel closes the start table itself (cell_start0[C] = n)
        ProfScope ps(ctx, "cells", (double)n);
        cells_mail = mail_open(MAIL_CELLS);
        if (key_embed) {
            mort0.reset(ctx, mort_len(n));
            segment_runs(ctx, CellKeyEmbedFn<KeyT>{keys0.get(), kp.pose_bits + fw, fw, reinterpret_cast<uint32_t*>(mort0.get())},
                         CellEmitFn{cell_key.get(), cell_start0.get()}, n, cellidx0.get(), d_total.get(), cell_start0.get(), cells_mail);
        } else {
            segment_runs(ctx, CellKeyFn<KeyT>{keys0.get(), kp.pose_bits}, CellEmitFn{cell_key.get(), cell_start0.get()}, n,
                         cellidx0.get(), d_total.get(), cell_start0.get(), cells_mail);
        }
    };
    if (key_embed) mort_r.release();
    if (key_bytes == 4)
        run(uint32_t{});
    else
        run(uint64_t{});
    perm0.swap(vals0);
    if (!key_embed) {
        mort0.reset(ctx, mort_len(n));
        ProfScope ps(ctx, "gather_morton", (double)n);
        if (mort32)
            gather_kernel<uint32_t><<<nblk(n), 256, 0, ctx.stream>>>(reinterpret_cast<uint32_t*>(mort0.get()),
                                                                     reinterpret_cast<const uint32_t*>(mort_r.get()), perm0.get(), n);
        else
            gather_kernel<uint64_t><<<nblk(n), 256, 0, ctx.stream>>>(mort0.get(), mort_r.get(), perm0.get(), n);
        OL_CHECK_LAUNCH();
    }
    mort_r.release();
    cells_pending = true;  // C arrives through the mailbox
    cp_valid = false;  // the (cell, pose) table is built on first use (ensure_cell_poses)
    built = true;
    base_dirty = any_dead;  // points removed before a rebuild are dropped from the base order lazily
}

// (cell, pose) pairs = runs of equal (cell, pose) in the base order (octree_manager.py:166-169).  Only the node counters
// and the cell exports read the table, so it is built on first use - but always before the base order loses points, because
// a pose keeps its octree in a cell after its points are filtered away.
void Forest::ensure_cell_poses() {
    if (cp_valid) return;
    const uint32_t n = A0;
    DevBuf<unsigned long long> d_total(ctx, 1);
    const GroupPoseKeyFn key{cellidx0.get(), perm0.get(), d_seg_start.get(), d_seg_pose.get(), (int)seg_pose.size()};
    const size_t cp_max = std::min<size_t>(n, (size_t)C * (size_t)std::max(n_poses, 1));
    DevBuf<uint32_t> cell_tmp(ctx, cp_max);
    DevBuf<int32_t> pose_tmp(ctx, cp_max);
    cell_first_pose.reset(ctx, C);
    {
        ProfScope ps(ctx, "cell_poses", (double)n);
        segment_runs(ctx, key, CellPoseEmitFn{cell_tmp.get(), pose_tmp.get(), cell_first_pose.get(), cellidx0.get()}, n, nullptr,
                     d_total.get());
    }
    CP = (uint32_t)read_u64(d_total.get());
    cp_cell.reset(ctx, CP);
    cp_pose.reset(ctx, CP);
    d2d(ctx, cp_cell.get(), cell_tmp.get(), CP);
    d2d(ctx, cp_pose.get(), pose_tmp.get(), CP);
    cp_valid = true;
}

// drop dead points from the base order (cells keep their index even when they become empty)
void Forest::compact_base() {
    if (!base_dirty) return;
    ensure_cell_poses();
    ensure_alive();
    const uint32_t n = A0;
    if (n) {
        CompactTables t;
        const uint32_t total = compact_tables(alive_r.get(), perm0.get(), n, t);
        DevBuf<uint32_t> p2(ctx, total), c2(ctx, total), s2(ctx, (size_t)C + 1);
        DevBuf<uint64_t> m2(ctx, mort_len(total));
        compact_move(t, n, perm0.get(), mort0.get(), cellidx0.get(), p2.get(), m2.get(), c2.get());
        {
            ProfScope ps(ctx, "compact");
            remap_starts_kernel<<<nblk((size_t)C + 1), 256, 0, ctx.stream>>>(cell_start0.get(), t.bits.get(), t.word_off.get(), C, n, total,
                                                                              s2.get());
            OL_CHECK_LAUNCH();
        }
        perm0.swap(p2);
        mort0.swap(m2);
        cellidx0.swap(c2);
        cell_start0.swap(s2);
        A0 = total;
    }
    base_dirty = false;
}

// keep flags (keep[via ? via[i] : i]) -> bit words, tile offsets; returns the number of kept positions
uint32_t Forest::compact_tables(const uint8_t* keep, const uint32_t* via, uint32_t n, CompactTables& t) {
    const uint32_t tiles = (n + CMP_TILE - 1) / CMP_TILE;
    t.bits.reset(ctx, (n + 31) / 32);
    t.word_off.reset(ctx, (n + 31) / 32);
    t.tile_off.reset(ctx, tiles);
    DevBuf<unsigned long long> d_total(ctx, 1);
    {
        ProfScope ps(ctx, "compact", (double)n);
        keep_bits_kernel<<<tiles, CMP_THREADS, 0, ctx.stream>>>(keep, via, n, t.bits.get(), t.tile_off.get());
        OL_CHECK_LAUNCH();
    }
    const Mail mail = mail_open(MAIL_COMPACT);
    transform_scan<uint32_t>(ctx, ScanPtrIn<uint32_t>{t.tile_off.get()}, ScanPtrOut<uint32_t>{t.tile_off.get()}, tiles, d_total.get(), "scan",
                             mail);
    const MailResult r = mail_take(mail);  // the error word rides along (RANSAC flags of the launch before)
    note_ransac_flags(r.err);
    return (uint32_t)r.total;
}

void Forest::compact_move(CompactTables& t, uint32_t n, const uint32_t* perm_in, const uint64_t* mort_in, const uint32_t* aux_in,
                          uint32_t* perm_out, uint64_t* mort_out, uint32_t* aux_out) {
    const uint32_t tiles = (n + CMP_TILE - 1) / CMP_TILE;
    ProfScope ps(ctx, "compact", (double)n);
    if (mort32)
        compact_move_kernel<uint32_t><<<tiles, CMP_THREADS, 0, ctx.stream>>>(t.bits.get(), t.tile_off.get(), n, perm_in,
                                                                             reinterpret_cast<const uint32_t*>(mort_in), aux_in, perm_out,
                                                                             reinterpret_cast<uint32_t*>(mort_out), aux_out, t.word_off.get());
    else
        compact_move_kernel<uint64_t><<<tiles, CMP_THREADS, 0, ctx.stream>>>(t.bits.get(), t.tile_off.get(), n, perm_in, mort_in, aux_in,
                                                                             perm_out, mort_out, aux_out, t.word_off.get());
    OL_CHECK_LAUNCH();
}

// The byte map of stored points is only read when the base order is rebuilt or compacted (insert after a removal,
// a second subdivide): apply_keep leaves it stale and it is re-derived here from the current order, whose positions
// are exactly the stored points - |kept| scattered byte stores instead of |removed| ones inside every apply_keep.
void Forest::ensure_alive() {
    if (!alive_stale) return;
    alive_stale = false;
    if (!alive_r.get()) alive_r.reset(ctx, cap);
    OL_CUDA(cudaMemsetAsync(alive_r.get(), 0, N, ctx.stream));
    if (cap > N) OL_CUDA(cudaMemsetAsync(alive_r.get() + N, 1, cap - N, ctx.stream));
    if (A) {
        ProfScope ps(ctx, "compact", (double)A);
        mark_alive_kernel<<<nblk(A), 256, 0, ctx.stream>>>(perm.get(), A, alive_r.get());
        OL_CHECK_LAUNCH();
    }
}

// The Morton codes carry MORTON_INITIAL_DEPTH levels at first (3 bits each); a subdivision that goes deeper
// recomputes them at the full depth for the base order and for the current order.
void Forest::extend_morton() {
    if (kp.depth >= max_depth) return;
    // in two stages: up to MORTON32_MAX_DEPTH levels the codes stay 32-bit words (a quarter less partition traffic)
    kp.depth = kp.depth < MORTON32_MAX_DEPTH ? std::min(max_depth, MORTON32_MAX_DEPTH) : max_depth;
    const bool was32 = mort32;
    mort32 = kp.depth <= MORTON32_MAX_DEPTH;
    ProfScope ps(ctx, "keygen", (double)(A0 + (shaped ? A : 0)));
    auto recompute = [&](DevBuf<uint64_t>& buf, const uint32_t* perm_p, const uint32_t* cell_of, const uint32_t* lcell_p, uint32_t n) {
        if (was32 != mort32) buf.reset(ctx, mort_len(n));
        if (n == 0) return;
        if (mort32)
            remorton_kernel<uint32_t><<<nblk(n), 256, 0, ctx.stream>>>(P64.get(), perm_p, cell_of, lcell_p, cell_key.get(), kp, n,
                                                                       reinterpret_cast<uint32_t*>(buf.get()));
        else
            remorton_kernel<uint64_t><<<nblk(n), 256, 0, ctx.stream>>>(P64.get(), perm_p, cell_of, lcell_p, cell_key.get(), kp, n, buf.get());
        OL_CHECK_LAUNCH();
    };
    recompute(mort0, perm0.get(), cellidx0.get(), nullptr, A0);
    if (shaped) recompute(mort, perm.get(), leaf_of.get(), lcell.get(), A);
}

// current shape := one leaf per cell
void Forest::reset_shape() {
    materialize_snapshot();
    build();
    compact_base();
    A = A0;
    A_shape = A0;
    L = C;
    I = 0;
    depth_reached = 0;
    // the current point order starts as the base order: not copied until somebody needs it as such - the first
    // partition level reads the base arrays directly (split_levels), everything else calls materialize_order()
    perm.reset(ctx, 0);
    mort.reset(ctx, 0);
    leaf_of.reset(ctx, 0);
    order_virgin = true;
    lstart.reset(ctx, (size_t)L + 1);
    d2d(ctx, lstart.get(), cell_start0.get(), (size_t)L + 1);
    lcell.reset(ctx, L);
    lparent.reset(ctx, L);
    lpath.reset(ctx, L);
    ldepth.reset(ctx, L);
    lchild.reset(ctx, L);
    if (L) {
        init_leaves_kernel<<<nblk(L), 256, 0, ctx.stream>>>(L, lcell.get(), lparent.get(), lpath.get(), ldepth.get(), lchild.get());
        OL_CHECK_LAUNCH();
    }
    level_ibegin.clear();
    shaped = true;
    order_valid = blocks_valid = ransac_valid = false;
}

void Forest::materialize_order() {
    if (!order_virgin) return;
    order_virgin = false;
    perm.reset(ctx, A);
    mort.reset(ctx, mort_len(A));
    leaf_of.reset(ctx, A);
    d2d(ctx, perm.get(), perm0.get(), A);
    d2d(ctx, mort.get(), mort0.get(), mort_len(A));
    d2d(ctx, leaf_of.get(), cellidx0.get(), A);
}

void Forest::ensure_shape() {
    if (shaped) return;
    reset_shape();
    if (replay_pending) replay_shape();
    materialize_order();
}

void Forest::save_shape() {
    if (replay_pending) return;  // still waiting for a rebuild: the recorded shape is the current one
    OL_REQUIRE(depth_reached <= REPLAY_MAX_DEPTH, OL_ERR_STATE,
               "keeping the shape of a grid subdivided deeper than " + std::to_string(REPLAY_MAX_DEPTH) +
                   " levels across a rebuild (a pose inserted after the subdivision, or a second subdivide) is not supported");
    sp_n = I;
    sp_q.reset(ctx, (size_t)I * 3);
    sp_depth.reset(ctx, I);
    sp_path.reset(ctx, I);
    save_shape_kernel<<<nblk(I), 256, 0, ctx.stream>>>(I, icell.get(), idepth.get(), ipath.get(), cell_key.get(), kp, sp_q.get(),
                                                       sp_depth.get(), sp_path.get());
    OL_CHECK_LAUNCH();
    if (epochs_valid) {
        sp_epoch.reset(ctx, I);
        d2d(ctx, sp_epoch.get(), iepoch.get(), I);
    } else {
        sp_epoch.release();  // every recorded node has epoch 1
    }
    replay_pending = true;
}

// The shape as data: the split (internal) nodes by cell coordinates, depth and Morton path.  `Octree.subdivide_as`
// (octree.py:34-53, 222-227) copies one octree's subdivision scheme onto another: export the one, impose it on the other.
uint32_t Forest::export_shape(long long* q_host, uint32_t* depth_host, unsigned long long* path_host) {
    ensure_shape();
    if (!q_host || I == 0) return I;
    OL_REQUIRE(depth_reached <= REPLAY_MAX_DEPTH, OL_ERR_STATE,
               "exporting the shape of a tree deeper than " + std::to_string(REPLAY_MAX_DEPTH) + " levels is not supported");
    DevBuf<long long> q(ctx, (size_t)I * 3);
    DevBuf<uint32_t> d(ctx, I);
    DevBuf<uint64_t> p(ctx, I);
    save_shape_kernel<<<nblk(I), 256, 0, ctx.stream>>>(I, icell.get(), idepth.get(), ipath.get(), cell_key.get(), kp, q.get(), d.get(), p.get());
    OL_CHECK_LAUNCH();
    d2h(ctx, q_host, q.get(), (size_t)I * 3);
    d2h(ctx, depth_host, d.get(), I);
    d2h(ctx, reinterpret_cast<uint64_t*>(path_host), p.get(), I);
    ctx.sync();
    return I;
}

// The next time the shape is needed it is rebuilt from the cell roots and exactly the listed nodes are split (where the
// forest has them: cells it does not hold, and nodes below an unlisted parent, are ignored); n = 0 collapses everything.
void Forest::impose_shape(const long long* q_host, const uint32_t* depth_host, const unsigned long long* path_host, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i)
        OL_REQUIRE(depth_host[i] < (uint32_t)REPLAY_MAX_DEPTH, OL_ERR_INVALID,
                   "a shape deeper than " + std::to_string(REPLAY_MAX_DEPTH) + " levels cannot be imposed");
    materialize_snapshot();
    sp_n = n;
    sp_q.reset(ctx, (size_t)std::max<uint32_t>(n, 1) * 3);
    sp_depth.reset(ctx, std::max<uint32_t>(n, 1));
    sp_path.reset(ctx, std::max<uint32_t>(n, 1));
    sp_epoch.release();
    if (n) {
        h2d(ctx, sp_q.get(), q_host, (size_t)n * 3);
        h2d(ctx, sp_depth.get(), depth_host, n);
        h2d(ctx, sp_path.get(), reinterpret_cast<const uint64_t*>(path_host), n);
        ctx.sync();  // the host arrays belong to the caller
    }
    replay_pending = true;
    shaped = false;
    epochs_valid = false;
    order_valid = blocks_valid = ransac_valid = false;
}

void Forest::assign_epochs(int epoch, const uint64_t* sorted_keys, const uint32_t* sorted_vals, uint32_t n_saved) {
    iepoch.reset(ctx, I);
    if (I) {
        assign_epochs_kernel<<<nblk(I), 256, 0, ctx.stream>>>(I, icell.get(), idepth.get(), ipath.get(), sorted_keys, sorted_vals,
                                                              sp_epoch.get(), n_saved, (uint32_t)epoch, iepoch.get());
        OL_CHECK_LAUNCH();
    }
    epochs_valid = true;
}

void Forest::replay_shape() {
    replay_pending = false;
    const uint32_t n = sp_n;
    if (n == 0 || L == 0 || A == 0) return;
    DevBuf<uint64_t> k0(ctx, n), k1(ctx, n);
    DevBuf<uint32_t> v0(ctx, n), v1(ctx, n);
    replay_keys_kernel<<<nblk(n), 256, 0, ctx.stream>>>(n, sp_q.get(), sp_depth.get(), sp_path.get(), cell_key.get(), C, kp, k0.get(),
                                                        v0.get());
    OL_CHECK_LAUNCH();
    const int w = radix_sort_pairs<uint64_t>(ctx, k0.get(), k1.get(), v0.get(), v1.get(), n, 0, 64);
    split_levels(nullptr, nullptr, nullptr, 0, w ? k1.get() : k0.get(), n);
    if (n_subdivide_calls >= 2) assign_epochs(n_subdivide_calls, w ? k1.get() : k0.get(), w ? v1.get() : v0.get(), n);
    sp_q.release();
    sp_depth.release();
    sp_path.release();
    sp_epoch.release();
    sp_n = 0;
}

// ---------------------------------------------------------------------------------------------
// K4: Grid.subdivide (grid.py:244-258) -> OctreeManager.subdivide (octree_manager.py:36-66).
// Every call rebuilds the shape from the cell roots, as the reference's fresh scheme octree does.
// ---------------------------------------------------------------------------------------------
void Forest::subdivide(const SplitRule& rule, const int32_t* poses, int n_listed) {
    OL_REQUIRE(!rule.first_level.empty() && rule.first_level[0] == 0, OL_ERR_INVALID, "split rule must start at level 0");
    for (size_t e = 1; e < rule.first_level.size(); ++e)
        OL_REQUIRE(rule.first_level[e] > rule.first_level[e - 1], OL_ERR_INVALID, "split rule levels must ascend");
    DevBuf<uint8_t> listed, tables;
    if (n_listed > 0) {
        std::vector<uint8_t> h(std::max(n_poses, 1), 0);
        for (int i = 0; i < n_listed; ++i) {
            OL_REQUIRE(poses[i] >= 0 && poses[i] < n_poses, OL_ERR_POSE, "unknown pose index " + std::to_string(poses[i]));
            h[poses[i]] = 1;
        }
        listed.reset(ctx, h.size());
        h2d(ctx, listed.get(), h.data(), h.size());  // pageable source: staged before the call returns
    }
    if (rule.tables_host) {
        OL_REQUIRE(rule.table_len > 0, OL_ERR_INVALID, "empty split table");
        const size_t bytes = (size_t)rule.table_len * rule.first_level.size();
        tables.reset(ctx, bytes);
        h2d(ctx, tables.get(), rule.tables_host, bytes);
    }
    // From the second call on the reference's leaf order depends on WHEN a node was split (forest.cuh): keep the current
    // shape (with its epochs) as the record the new shape's nodes are matched against.
    const int epoch = n_subdivide_calls + 1;
    bool have_prev = false;
    if (epoch >= 2) {
        if (!replay_pending && shaped && I > 0) save_shape();  // a pending replay already holds the shape it stands for
        have_prev = replay_pending && sp_n > 0;
    }
    replay_pending = false;  // a fresh scheme replaces whatever shape was recorded
    reset_shape();  // after the argument checks: the deferred copy of the base order must not outlive an early error
    n_subdivide_calls = epoch;
    if (L == 0 || A == 0) {
        materialize_order();
    } else {
        split_levels(&rule, tables.get(), listed.get(), n_listed, nullptr, 0);
    }
    if (epoch >= 2) {
        if (have_prev && I > 0) {
            const uint32_t n = sp_n;
            DevBuf<uint64_t> k0(ctx, n), k1(ctx, n);
            DevBuf<uint32_t> v0(ctx, n), v1(ctx, n);
            replay_keys_kernel<<<nblk(n), 256, 0, ctx.stream>>>(n, sp_q.get(), sp_depth.get(), sp_path.get(), cell_key.get(), C, kp,
                                                                k0.get(), v0.get());
            OL_CHECK_LAUNCH();
            const int w = radix_sort_pairs<uint64_t>(ctx, k0.get(), k1.get(), v0.get(), v1.get(), n, 0, 64);
            assign_epochs(epoch, w ? k1.get() : k0.get(), w ? v1.get() : v0.get(), n);
        } else {
            assign_epochs(epoch, nullptr, nullptr, 0);
        }
        sp_q.release();
        sp_depth.release();
        sp_path.release();
        sp_epoch.release();
        sp_n = 0;
    } else {
        epochs_valid = false;
    }
    static const bool no_prefetch = getenv("OL_NO_PREFETCH") != nullptr;  // debug: build the derived tables on demand only
    if (!no_prefetch) prefetch_tables();
}

// the internal-node arrays grow geometrically, so a level appends in place
void Forest::reserve_internal(size_t need) {
    if (need <= icap && istart.get()) return;
    const size_t ncap = std::max<size_t>(need + need / 2, std::max<size_t>(icap * 4, 65536));  // few re-allocations: each is six copies
    DevBuf<uint32_t> s2(ctx, ncap), c2(ctx, ncap);
    DevBuf<uint8_t> d2(ctx, ncap), ch2(ctx, ncap);
    DevBuf<uint64_t> p2(ctx, ncap);
    DevBuf<int32_t> pa2(ctx, ncap);
    if (I) {
        d2d(ctx, s2.get(), istart.get(), I);
        d2d(ctx, c2.get(), icell.get(), I);
        d2d(ctx, d2.get(), idepth.get(), I);
        d2d(ctx, p2.get(), ipath.get(), I);
        d2d(ctx, pa2.get(), iparent.get(), I);
        d2d(ctx, ch2.get(), ichild.get(), I);
    }
    istart.swap(s2);
    icell.swap(c2);
    idepth.swap(d2);
    ipath.swap(p2);
    iparent.swap(pa2);
    ichild.swap(ch2);
    icap = ncap;
}

// The level loop of K4.  Decision per leaf of the current level: count criterion (all points, or the points of
// the listed poses), count table, or - replay_keys != nullptr - membership in a recorded shape.
// Per level: [weighted count] -> decide + scan (ONE kernel: transform_scan) -> read-back of (number of splits, error
// word) -> digit histograms -> ONE scan of the flat histogram buffer -> new tables + deltas -> fused rank-and-move.
void Forest::split_levels(const SplitRule* rule, const uint8_t* d_tables, const uint8_t* d_listed, int n_listed,
                          const uint64_t* replay_keys, uint32_t n_replay) {
    const int S = (int)seg_pose.size();
    // an error thrown while the current order is still the un-copied base order: drop the shape, the next call rebuilds it
    struct VirginGuard {
        Forest* f;
        int live = std::uncaught_exceptions();
        ~VirginGuard() {
            if (f->order_virgin && std::uncaught_exceptions() > live) {
                f->order_virgin = false;
                f->shaped = false;
            }
        }
    } virgin_guard{this};
    DevBuf<unsigned long long> d_tot(ctx, 2);
    const uint32_t tiles = (A + PART_TILE - 1) / PART_TILE;
    DevBuf<uint32_t> perm_b(ctx, A), leaf_b(ctx, A);
    DevBuf<uint64_t> mort_b(ctx, mort_len(A));
    level_ibegin.assign(1, 0u);
    // Upper bound on the number of leaves one level can split: a leaf only splits when it holds more points than the
    // smallest count the rule accepts.  It sizes the buffers of the SPECULATIVE histogram pass below.
    auto split_bound = [&](int level) -> size_t {
        size_t per_leaf = 1;  // smallest point count of a splitting leaf
        if (!rule) return std::min<size_t>(L, n_replay);
        const int e = rule->entry_for(level);
        if (rule->tables_host) {
            const uint8_t* t = rule->tables_host + (size_t)e * (size_t)rule->table_len;
            int64_t c = 0;
            while (c < rule->table_len && !t[c]) ++c;
            per_leaf = (size_t)std::max<int64_t>(c, 0);
        } else {
            per_leaf = rule->max_points[e] >= (int64_t)A ? (size_t)A + 1 : (size_t)std::max<int64_t>(rule->max_points[e] + 1, 0);
        }
        return per_leaf == 0 ? (size_t)L : std::min<size_t>(L, (size_t)A / per_leaf);
    };
    constexpr size_t SPEC_MAX_LEAVES = (size_t)4 << 20;  // 128 MB of speculative counters at most
    // internal-node arrays: room for two full levels of splits at once, so that a typical run never re-allocates (each
    // re-allocation is six device-to-device copies)
    if (rule && !rule->tables_host && L && A) reserve_internal(std::min<size_t>((size_t)I + 2 * std::max<size_t>(split_bound(0) , A / ((size_t)std::max<int64_t>(rule->max_points[0], 0) + 1)), (size_t)8 << 20));
    for (int level = 0;; ++level) {
        // level 0 of a fresh shape partitions straight out of the base order (reset_shape made no copy)
        const uint32_t* src_leaf = order_virgin ? cellidx0.get() : leaf_of.get();
        const uint32_t* src_perm = order_virgin ? perm0.get() : perm.get();
        DevBuf<uint32_t> sinfo(ctx, L), wcount;
        if (n_listed > 0) {
            wcount.reset(ctx, L);
            wcount.zero();
            {
                ProfScope ps(ctx, "weighted_count", (double)A);
                weighted_count_kernel<<<nblk(A), 256, 0, ctx.stream>>>(src_leaf, src_perm, ldepth.get(), level,
                                                                       d_seg_start.get(), d_seg_pose.get(), S, d_listed, A,
                                                                       wcount.get());
                OL_CHECK_LAUNCH();
            }
        }
        DecideIn<DECIDE_THRESHOLD, false> din{};
        din.lstart = lstart.get();
        din.ldepth = ldepth.get();
        din.level = level;
        din.wcount = wcount.get();
        din.max_depth = max_depth;
        din.replay_keys = replay_keys;
        din.n_replay = n_replay;
        din.lcell = lcell.get();
        din.lpath = lpath.get();
        din.err = d_err.get();
        int mode = replay_keys ? DECIDE_REPLAY : DECIDE_THRESHOLD;
        if (rule) {
            const int e = rule->entry_for(level);
            if (rule->tables_host) {
                din.table = d_tables + (size_t)e * (size_t)rule->table_len;
                din.table_len = rule->table_len;
                din.beyond = rule->beyond[e];
                mode = DECIDE_TABLE;
            } else {
                din.max_points = rule->max_points[e];
            }
        }
        // decide + scan; the number of splitting leaves (and the error word) is POSTED to the host by the kernel
        const Mail mail = mail_open(MAIL_LEVEL);
        auto run_decide = [&](auto tag) {  // every instantiation has the same fields
            decltype(tag) d{};
            static_assert(sizeof(d) == sizeof(din), "DecideIn instantiations must share one layout");
            memcpy(&d, &din, sizeof(din));
            transform_scan<uint32_t>(ctx, d, DecideOut{sinfo.get()}, L, d_tot.get(), "part_decide", mail);
        };
        const bool cap = level >= max_depth;
        if (mode == DECIDE_REPLAY)
            run_decide(DecideIn<DECIDE_REPLAY, false>{});  // a recorded shape never reaches the depth cap
        else if (mode == DECIDE_TABLE)
            cap ? run_decide(DecideIn<DECIDE_TABLE, true>{}) : run_decide(DecideIn<DECIDE_TABLE, false>{});
        else
            cap ? run_decide(DecideIn<DECIDE_THRESHOLD, true>{}) : run_decide(DecideIn<DECIDE_THRESHOLD, false>{});
        // Speculation: the digit histograms of the level only need the decisions (device side), not their number, so the
        // pass is enqueued BEHIND the decision kernel with counters sized by the bound - the host reads the mailbox while
        // it runs and the GPU never waits for the host.  A level that splits nothing makes the pass return at once.
        const size_t bound = std::max<size_t>(split_bound(level), 1);
        const bool speculate = level < kp.depth && bound <= SPEC_MAX_LEAVES;
        DevBuf<uint32_t> tile_hist(ctx, (size_t)8 * tiles), leaf_cnt;
        auto launch_hist = [&](size_t stride, const unsigned long long* d_nsplit) {
            const uint64_t* sm = order_virgin ? mort0.get() : mort.get();
            const int shift = 3 * (kp.depth - 1 - level);
            OL_CUDA(cudaMemsetAsync(leaf_cnt.get(), 0, (size_t)8 * stride * 4, ctx.stream));
            ProfScope ps(ctx, "part_hist", (double)A);
            if (mort32)
                part_hist_kernel<uint32_t><<<tiles, PART_THREADS, 0, ctx.stream>>>(src_leaf, reinterpret_cast<const uint32_t*>(sm), sinfo.get(),
                                                                                   A, tiles, (uint32_t)stride, shift, tile_hist.get(),
                                                                                   leaf_cnt.get(), d_nsplit);
            else
                part_hist_kernel<uint64_t><<<tiles, PART_THREADS, 0, ctx.stream>>>(src_leaf, sm, sinfo.get(), A, tiles, (uint32_t)stride, shift,
                                                                                   tile_hist.get(), leaf_cnt.get(), d_nsplit);
            OL_CHECK_LAUNCH();
        };
        if (speculate) {
            leaf_cnt.reset(ctx, (size_t)8 * bound);
            launch_hist(bound, d_tot.get());
        }
        const MailResult res = mail_take(mail);  // the level's ONE host wait (no stream synchronisation)
        throw_device_errors(res.err);  // depth cap of this level, out-of-node points met by the previous level's move
        const uint32_t n_split = (uint32_t)res.total;
        if (n_split == 0) break;
        const uint32_t L_new = L + 7u * n_split;
        OL_REQUIRE((unsigned long long)I + n_split < (1ull << 29), OL_ERR_RANGE, "too many internal nodes");
        OL_REQUIRE(!speculate || n_split <= bound, OL_ERR_INTERNAL, "split bound violated");
        depth_reached = level + 1;
        size_t stride = bound;
        if (!speculate) {
            if (level >= kp.depth) {
                materialize_order();
                src_leaf = leaf_of.get();
                src_perm = perm.get();
                extend_morton();
                mort_b.reset(ctx, mort_len(A));  // the word size may have changed
            }
            stride = n_split;
            leaf_cnt.reset(ctx, (size_t)8 * stride);
            launch_hist(stride, nullptr);
        }
        const uint64_t* src_mort = order_virgin ? mort0.get() : mort.get();
        const int shift_now = 3 * (kp.depth - 1 - level);  // kp.depth may have grown (extend_morton)
        // ONE scan over [ tile_hist (8 x tiles) | leaf_cnt (8 x n_split, gathered out of the strided counters) ]
        const size_t flat_len = (size_t)8 * tiles + (size_t)8 * n_split;
        DevBuf<uint32_t> hist(ctx, flat_len), delta(ctx, (size_t)n_split * 8);
        transform_scan<uint32_t>(ctx, HistIn{tile_hist.get(), leaf_cnt.get(), 8u * tiles, n_split, (uint32_t)stride},
                                 ScanPtrOut<uint32_t>{hist.get()}, flat_len, d_tot.get() + 1);
        // new leaf / internal tables + partition deltas
        reserve_internal((size_t)I + n_split);
        DevBuf<uint32_t> lstart_n(ctx, (size_t)L_new + 1), lcell_n(ctx, L_new);
        DevBuf<int32_t> lparent_n(ctx, L_new);
        DevBuf<uint64_t> lpath_n(ctx, L_new);
        DevBuf<uint8_t> ldepth_n(ctx, L_new), lchild_n(ctx, L_new);
        {
            ProfScope ps(ctx, "part_expand");
            expand_leaves_kernel<<<nblk(L), 256, 0, ctx.stream>>>(L, A, I, sinfo.get(), hist.get(), d_tot.get() + 1, tiles, n_split,
                                                                  lstart.get(), lcell.get(), lparent.get(), lpath.get(), ldepth.get(),
                                                                  lchild.get(), L_new, lstart_n.get(), lcell_n.get(), lparent_n.get(),
                                                                  lpath_n.get(), ldepth_n.get(), lchild_n.get(), istart.get(),
                                                                  icell.get(), idepth.get(), ipath.get(), iparent.get(), ichild.get(),
                                                                  delta.get());
            OL_CHECK_LAUNCH();
        }
        {
            ProfScope ps(ctx, "part_move", (double)A);
            if (mort32)
                part_move_kernel<uint32_t><<<tiles, PART_THREADS, 0, ctx.stream>>>(
                    src_leaf, reinterpret_cast<const uint32_t*>(src_mort), src_perm, sinfo.get(), hist.get(), delta.get(), A, tiles,
                    shift_now, level, leaf_b.get(), reinterpret_cast<uint32_t*>(mort_b.get()), perm_b.get(), P64.get(), lcell.get(),
                    cell_key.get(), kp, d_err.get());
            else
                part_move_kernel<uint64_t><<<tiles, PART_THREADS, 0, ctx.stream>>>(
                    src_leaf, src_mort, src_perm, sinfo.get(), hist.get(), delta.get(), A, tiles, shift_now, level, leaf_b.get(),
                    mort_b.get(), perm_b.get(), P64.get(), lcell.get(), cell_key.get(), kp, d_err.get());
            OL_CHECK_LAUNCH();
        }
        lstart.swap(lstart_n);
        lcell.swap(lcell_n);
        lparent.swap(lparent_n);
        lpath.swap(lpath_n);
        ldepth.swap(ldepth_n);
        lchild.swap(lchild_n);
        perm.swap(perm_b);
        mort.swap(mort_b);
        leaf_of.swap(leaf_b);
        if (order_virgin) {  // the swapped-out buffers were the empty placeholders
            order_virgin = false;
            perm_b.reset(ctx, A);
            leaf_b.reset(ctx, A);
            mort_b.reset(ctx, mort_len(A));
        }
        L = L_new;
        I += n_split;
        level_ibegin.push_back(I);
    }
    materialize_order();  // no level split anything: the current order is the base order
    order_valid = blocks_valid = ransac_valid = false;
}

// ---------------------------------------------------------------------------------------------
// K5: leaf enumeration order of the reference + leaf geometry
//   order inside a cell = for every internal node in DFS pre-order, its leaf children by child id
//   (octree_base.py:48-49 appends, octree.py:183-191 removes the parent and appends 8 children);
//   an unsplit root is the cell's only leaf.  DFS pre-order of the internal nodes = ascending
//   (range start, depth).
// ---------------------------------------------------------------------------------------------
void Forest::ensure_order() {
    ensure_shape();
    if (order_valid) return;
    cache_rank.reset(ctx, L);
    leaf_by_cache.reset(ctx, L);
    leaf_corner.reset(ctx, (size_t)L * 3);
    leaf_edge.reset(ctx, L);
    cell_leaf_begin.reset(ctx, (size_t)C + 1);
    if (L == 0) {
        cell_leaf_begin.zero();
        order_valid = true;
        return;
    }
    ProfScope ps(ctx, "leaf_order", (double)L);
    DevBuf<uint32_t> irank(ctx, I), imask(ctx, I), nlc_r(ctx, I), leafbase(ctx, I), cell_ifirst(ctx, C);
    if (I) {
        OL_REQUIRE((int)level_ibegin.size() == depth_reached + 1 && level_ibegin.back() == I, OL_ERR_INTERNAL,
                   "internal-node level table out of step with the shape");
        LevelBegins lv{};
        lv.n = depth_reached;
        for (int d = 0; d <= depth_reached; ++d) lv.b[d] = level_ibegin[d];
        imask.zero();
        internal_mask_kernel<<<nblk(I), 256, 0, ctx.stream>>>(I, iparent.get(), ichild.get(), imask.get());
        OL_CHECK_LAUNCH();
        internal_rank_kernel<<<nblk(I), 256, 0, ctx.stream>>>(I, idepth.get(), icell.get(), ipath.get(), imask.get(), lv, irank.get(),
                                                              nlc_r.get(), cell_ifirst.get());
        OL_CHECK_LAUNCH();
        exclusive_scan_u32(ctx, nlc_r.get(), leafbase.get(), I, nullptr);
    }
    cell_first_leaf_kernel<<<nblk(L), 256, 0, ctx.stream>>>(L, lcell.get(), cell_leaf_begin.get());
    OL_CHECK_LAUNCH();
    leaf_order_geometry_kernel<<<nblk(L), 256, 0, ctx.stream>>>(L, lcell.get(), lparent.get(), lchild.get(), lpath.get(), ldepth.get(),
                                                                irank.get(), imask.get(), leafbase.get(), cell_ifirst.get(),
                                                                cell_leaf_begin.get(), cell_key.get(), kp, cache_rank.get(),
                                                                leaf_by_cache.get(), leaf_corner.get(), leaf_edge.get());
    OL_CHECK_LAUNCH();
    OL_CUDA(cudaMemcpyAsync(cell_leaf_begin.get() + C, &L, 4, cudaMemcpyHostToDevice, ctx.stream));
    order_valid = true;
}

// ---------------------------------------------------------------------------------------------
// non-empty (pose, leaf) blocks = maximal runs of equal (leaf, pose) in the current point order
// ---------------------------------------------------------------------------------------------
void Forest::ensure_blocks() {
    ensure_blocks_enqueue();
    resolve_blocks();
}

void Forest::ensure_blocks_enqueue() {
    ensure_shape();
    if (blocks_valid) return;
    const int S = (int)seg_pose.size();
    NB = 0;
    max_block = 0;
    max_block_known = true;
    blocks_pending = false;
    blk_of_pos.reset(ctx, A);
    if (A == 0) {
        blk_start.reset(ctx, 1);
        blk_start.zero();
        blk_leaf.reset(ctx, 0);
        blk_pose.reset(ctx, 0);
        blocks_valid = true;
        return;
    }
    // ONE pass (primitives.cuh: runs_fused_kernel); the tables are sized by the upper bound min(A, L x P).  The kernel
    // posts the number of blocks to the host (read when somebody needs it: resolve_blocks) and closes the start table
    // itself (blk_start[NB] = A).
    const size_t nb_max = std::min<size_t>(A, (size_t)L * (size_t)std::max(n_poses, 1));
    d_nb.reset(ctx, 1);
    blk_start.reset(ctx, nb_max + 1);
    blk_leaf.reset(ctx, nb_max);
    blk_pose.reset(ctx, nb_max);
    const GroupPoseKeyFn key{leaf_of.get(), perm.get(), d_seg_start.get(), d_seg_pose.get(), S};
    {
        ProfScope ps(ctx, "blocks", (double)A);
        blocks_mail = mail_open(MAIL_BLOCKS);
        segment_runs(ctx, key, BlockEmitFn{blk_start.get(), blk_leaf.get(), blk_pose.get()}, A, blk_of_pos.get(), d_nb.get(),
                     blk_start.get(), blocks_mail);
    }
    blocks_pending = true;
    // the largest block is only computed when somebody needs it (ensure_max_block; the RANSAC batch layout folds it into
    // a pass it makes over the block sizes anyway and reads it together with its work size)
    d_max_block.reset(ctx, 1);
    d_max_block.zero();
    max_block_known = false;
    max_block_enqueued = false;
    blocks_valid = true;
}

void Forest::enqueue_max_block() {
    if (max_block_known || max_block_enqueued) return;
    ProfScope ps(ctx, "blocks");
    block_max_kernel<<<(unsigned)ctx.num_sms * 8, 256, 0, ctx.stream>>>(blk_start.get(), d_nb.get(), d_max_block.get());
    OL_CHECK_LAUNCH();
    max_block_enqueued = true;
}

// Enqueued at the end of a subdivide: nearly every next operation (RANSAC, filter, get_leaf_points, the counters) starts
// from the leaf order and the block table, and the host language needs tens of microseconds between two calls - the GPU
// builds the tables meanwhile instead of idling.  Nothing is waited for.
void Forest::prefetch_tables() {
    if (!shaped || A == 0 || L == 0) return;
    ensure_order();
    ensure_blocks_enqueue();
}

void Forest::ensure_max_block() {
    if (max_block_known) return;
    enqueue_max_block();
    max_block = read_u32(d_max_block.get());
    max_block_known = true;
}

// ---------------------------------------------------------------------------------------------
// K7: drop the positions whose keep flag is 0 (filter, RANSAC mask).  Tree shape unchanged.
// ---------------------------------------------------------------------------------------------
void Forest::apply_keep(const uint8_t* keep_pos) {
    const uint32_t n = A;
    if (n == 0) return;
    CompactTables t;
    const uint32_t total = compact_tables(keep_pos, nullptr, n, t);
    if (total == n) return;
    DevBuf<uint32_t> p2(ctx, total), l2(ctx, total), s2(ctx, (size_t)L + 1);
    DevBuf<uint64_t> m2(ctx, mort_len(total));
    compact_move(t, n, perm.get(), mort.get(), leaf_of.get(), p2.get(), m2.get(), l2.get());
    {
        ProfScope ps(ctx, "compact");
        remap_starts_kernel<<<nblk((size_t)L + 1), 256, 0, ctx.stream>>>(lstart.get(), t.bits.get(), t.word_off.get(), L, n, total, s2.get());
        OL_CHECK_LAUNCH();
    }
    perm.swap(p2);
    mort.swap(m2);
    leaf_of.swap(l2);
    lstart.swap(s2);
    A = total;
    any_dead = true;
    alive_stale = true;  // the byte map of stored points is re-derived on demand (ensure_alive)
    base_dirty = true;
    blocks_valid = false;
    ransac_valid = false;
}

// Grid.filter (grid.py:260-267): per (pose, leaf) block, keep iff the folded criteria say so
void Forest::filter(const uint8_t* keep_table_host, int64_t table_len, const int32_t* poses, int n_listed) {
    OL_REQUIRE(table_len > 0, OL_ERR_INVALID, "empty keep table");
    ensure_blocks();
    if (NB == 0) return;
    DevBuf<uint8_t> listed, table(ctx, (size_t)table_len), keep_blk(ctx, NB), keep_pos(ctx, A);
    if (n_listed > 0) {
        std::vector<uint8_t> h(std::max(n_poses, 1), 0);
        for (int i = 0; i < n_listed; ++i) {
            OL_REQUIRE(poses[i] >= 0 && poses[i] < n_poses, OL_ERR_POSE, "unknown pose index " + std::to_string(poses[i]));
            h[poses[i]] = 1;
        }
        listed.reset(ctx, h.size());
        h2d(ctx, listed.get(), h.data(), h.size());
    }
    h2d(ctx, table.get(), keep_table_host, (size_t)table_len);
    block_keep_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, blk_start.get(), blk_pose.get(), listed.get(), table.get(),
                                                        table_len, keep_blk.get());
    OL_CHECK_LAUNCH();
    pos_keep_from_block_kernel<<<nblk(A), 256, 0, ctx.stream>>>(A, blk_of_pos.get(), keep_blk.get(), keep_pos.get());
    OL_CHECK_LAUNCH();
    apply_keep(keep_pos.get());
}

}  // namespace ol
