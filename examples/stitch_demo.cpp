// examples/stitch_demo.cpp -- the reference-style C++ call sequence over include/octvr.hpp.
//   stitch_demo --dat FILE                 load a "VRv11" template, print its shape (no GPU needed)
//   stitch_demo --config JSON W IN_W IN_H BLEND OUT.i420
//        build the template on the GPU, push/pop three synthetic frames through AsyncMultiMapper,
//        write the last output frame (standard I420) to OUT.i420
//   stitch_demo --fast JSON W IN_W IN_H OUT.nv12
//        vr::FastMapper: full-frame template (use_roi = false), one stitch_nv12 of synthetic NV12 frames on the device
#include "octvr.hpp"
#include <cstdio>
#include <cstring>
#include <memory>
#include <sstream>

// the CUDA runtime calls the --fast mode needs for its device frames (declared here so the demo builds without CUDA headers)
extern "C" {
int cudaMalloc(void** p, size_t n);
int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
int cudaDeviceSynchronize(void);
int cudaFree(void* p);
}

static uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

int main(int argc, char** argv)
{
    try {
        if (argc >= 3 && !strcmp(argv[1], "--dat")) {
            std::ifstream f(argv[2], std::ios::binary);
            vr::MapperTemplate mt(f);
            printf("out %dx%d inputs %zu\n", mt.out_size.width, mt.out_size.height, mt.num_inputs());
            for (size_t i = 0; i < mt.num_inputs(); i++) {
                auto in = mt.input((int)i);
                printf("roi %d %d %d %d seam %d\n", in.roi.x, in.roi.y, in.roi.width, in.roi.height, in.seam_mask != nullptr);
            }
            return 0;
        }
        if (argc >= 7 && !strcmp(argv[1], "--incremental")) {
            // stitch_demo --incremental TO TO_OPTS_JSON WIDTH OUT.dat FROM FROM_OPTS_JSON [FROM FROM_OPTS_JSON ...]
            // the reference's own sequence (apps/octvr/dump.cpp:98-127): constructor, add_input per camera, create_masks, dump(ofstream)
            vr::MapperTemplate mt(argv[2], argv[3], atoi(argv[4]), -1);
            for (int k = 6; k + 1 < argc; k += 2) mt.add_input(argv[k], argv[k + 1]);
            mt.create_masks();
            std::ofstream out(argv[5], std::ios::binary);
            mt.dump(out);
            printf("out %dx%d inputs %zu seams %zu\n", mt.out_size.width, mt.out_size.height, mt.inputs.size(), mt.seam_masks.size());
            return 0;
        }
        if (argc >= 8 && !strcmp(argv[1], "--config")) {
            std::ifstream f(argv[2]);
            std::stringstream ss; ss << f.rdbuf();
            const int W = atoi(argv[3]), iw = atoi(argv[4]), ih = atoi(argv[5]), blend = atoi(argv[6]);
            vr::MapperTemplate mt = vr::MapperTemplate::from_config(ss.str(), W);
            const int n = (int)mt.num_inputs(), H = mt.out_size.height;
            std::vector<vr::Size> sizes(n, vr::Size{ iw, ih });
            std::vector<const vr::MapperTemplate*> mts{ &mt };
            std::unique_ptr<vr::AsyncMultiMapper> am(vr::AsyncMultiMapper::New(mts, sizes, mt.out_size, { blend }, { 0 }, { vr::RectD{ 0, 0, 1, 1 } }));
            std::vector<std::vector<uint8_t>> in(n, std::vector<uint8_t>((size_t)iw * ih * 3 / 2));
            std::vector<uint8_t> out((size_t)W * H * 3 / 2);
            for (int k = 0; k < 3; k++) {
                for (int c = 0; c < n; c++)
                    for (size_t o = 0; o < in[c].size(); o++) in[c][o] = (uint8_t)(splitmix64((1234 + k) ^ ((uint64_t)c << 32) ^ o) & 0xFF);
                std::vector<vr::YUV> frames;
                for (int c = 0; c < n; c++) {
                    uint8_t* b = in[c].data();
                    frames.emplace_back(vr::Plane{ b, (size_t)iw, 1 }, vr::Plane{ b + (size_t)iw * ih, (size_t)iw / 2, 1 },
                                        vr::Plane{ b + (size_t)iw * ih + (size_t)(iw / 2) * (ih / 2), (size_t)iw / 2, 1 });
                }
                vr::YUV o(vr::Plane{ out.data(), (size_t)W, 1 }, vr::Plane{ out.data() + (size_t)W * H, (size_t)W / 2, 1 },
                          vr::Plane{ out.data() + (size_t)W * H + (size_t)(W / 2) * (H / 2), (size_t)W / 2, 1 });
                am->push(frames, o);
                am->pop();
            }
            FILE* fo = fopen(argv[7], "wb");
            fwrite(out.data(), 1, out.size(), fo);
            fclose(fo);
            printf("ok %dx%d fps %.1f\n", W, H, am->fps());
            return 0;
        }
        if (argc >= 7 && !strcmp(argv[1], "--fast")) {
            std::ifstream f(argv[2]);
            std::stringstream ss; ss << f.rdbuf();
            const int W = atoi(argv[3]), iw = atoi(argv[4]), ih = atoi(argv[5]);
            vr::MapperTemplate mt = vr::MapperTemplate::from_config(ss.str(), W, -1, /*use_roi=*/false, /*create_masks=*/false);
            const int n = (int)mt.num_inputs(), H = mt.out_size.height;
            vr::FastMapper fm(mt, std::vector<vr::Size>(n, vr::Size{ iw, ih }));
            const size_t in_bytes = (size_t)iw * (ih + ih / 2), out_bytes = (size_t)W * (H + H / 2);
            std::vector<const uint8_t*> d_in(n);
            std::vector<uint8_t> h(in_bytes);
            for (int c = 0; c < n; c++) {
                for (size_t o = 0; o < in_bytes; o++) h[o] = (uint8_t)(splitmix64(0xFA57ull ^ ((uint64_t)c << 32) ^ o) & 0xFF);
                void* d = nullptr;
                if (cudaMalloc(&d, in_bytes) || cudaMemcpy(d, h.data(), in_bytes, 1)) throw std::runtime_error("cudaMalloc / cudaMemcpy");
                d_in[c] = (const uint8_t*)d;
            }
            void* d_out = nullptr;
            if (cudaMalloc(&d_out, out_bytes)) throw std::runtime_error("cudaMalloc");
            fm.stitch_nv12(d_in, std::vector<size_t>(n, (size_t)iw), (uint8_t*)d_out, (size_t)W);
            std::vector<uint8_t> out(out_bytes);
            if (cudaDeviceSynchronize() || cudaMemcpy(out.data(), d_out, out_bytes, 2)) throw std::runtime_error("cudaMemcpy");
            for (auto p : d_in) cudaFree((void*)p);
            cudaFree(d_out);
            FILE* fo = fopen(argv[6], "wb");
            fwrite(out.data(), 1, out.size(), fo);
            fclose(fo);
            printf("ok fast %dx%d\n", W, H);
            return 0;
        }
        fprintf(stderr, "usage: see source\n");
        return 2;
    } catch (const std::string& s) { fprintf(stderr, "std::string: %s\n", s.c_str()); return 10; }
    catch (const vr::NotImplemented&) { fprintf(stderr, "NotImplemented\n"); return 11; }
    catch (const std::exception& e) { fprintf(stderr, "exception: %s\n", e.what()); return 12; }
}
