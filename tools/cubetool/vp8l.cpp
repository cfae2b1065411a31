// vp8l.cpp -- WebP lossless (VP8L) decoder and encoder; see vp8l.hpp for why it lives here.
//
// Written from the published "WebP Lossless Bitstream Specification" and the RIFF container
// description; no libwebp code is used.  Pixels are handled as 32-bit ARGB (0xAARRGGBB) as in the
// specification.
#include "vp8l.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <stdexcept>
#include <string>

namespace vp8l {
	namespace {
		[[noreturn]] void fail(const std::string& what) {
			throw std::runtime_error("WebP lossless: " + what);
		}

		inline uint32_t rd32(const uint8_t* p) {
			return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24);
		}

		// ------------------------------------------------------------------------------------------
		// bit I/O, least-significant bit first
		// ------------------------------------------------------------------------------------------
		struct BitReader {
			const uint8_t* p;
			std::size_t n, pos = 0;
			uint64_t val = 0;
			int bits = 0;
			bool overrun = false;
			BitReader(const uint8_t* data, std::size_t size) : p(data), n(size) {}
			void fill() {
				while(bits <= 56 && pos < n) {
					val |= static_cast<uint64_t>(p[pos++]) << bits;
					bits += 8;
				}
			}
			uint32_t peek(int nb) {
				if(bits < nb) fill();
				return static_cast<uint32_t>(val & ((1ull << nb) - 1));
			}
			void skip(int nb) {
				if(bits < nb) {
					overrun = true;
					val = 0;
					bits = 0;
					return;
				}
				val >>= nb;
				bits -= nb;
			}
			uint32_t read(int nb) {
				if(nb == 0) return 0;
				const uint32_t v = peek(nb);
				skip(nb);
				return v;
			}
		};

		struct BitWriter {
			std::vector<uint8_t> out;
			uint64_t acc = 0;
			int bits = 0;
			void put(uint32_t v, int nb) {
				acc |= static_cast<uint64_t>(v) << bits;
				bits += nb;
				while(bits >= 8) {
					out.push_back(static_cast<uint8_t>(acc));
					acc >>= 8;
					bits -= 8;
				}
			}
			void flush() {
				if(bits > 0) out.push_back(static_cast<uint8_t>(acc));
				acc = 0;
				bits = 0;
			}
		};

		// ------------------------------------------------------------------------------------------
		// canonical prefix codes
		// ------------------------------------------------------------------------------------------
		constexpr int kMaxCodeLen = 15;
		constexpr int kCodeLengthCodes = 19;
		const uint8_t kCodeLengthOrder[kCodeLengthCodes] = {17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};

		inline uint32_t reverseBits(uint32_t v, int n) {
			uint32_t r = 0;
			for(int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
			return r;
		}

		// Decoding table indexed by the next `width` stream bits; entry = (symbol << 4) | length.
		struct PrefixCode {
			int width = 0;          // 0: single symbol, no bits are read
			uint32_t single = 0;    // that symbol
			std::vector<uint32_t> table;

			void build(const std::vector<uint8_t>& lens) {
				int count[kMaxCodeLen + 1] = {0};
				int used = 0, maxLen = 0;
				for(std::size_t s = 0; s < lens.size(); ++s) {
					if(lens[s] > kMaxCodeLen) fail("code length out of range");
					if(lens[s]) {
						++count[lens[s]];
						++used;
						single = static_cast<uint32_t>(s);
						maxLen = std::max<int>(maxLen, lens[s]);
					}
				}
				if(used == 0) fail("prefix code without symbols");
				if(used == 1) {
					width = 0;
					return;
				}
				// completeness (Kraft equality): libwebp rejects anything else too
				uint64_t kraft = 0;
				for(int l = 1; l <= kMaxCodeLen; ++l) kraft += static_cast<uint64_t>(count[l]) << (kMaxCodeLen - l);
				if(kraft != (1ull << kMaxCodeLen)) fail("incomplete or over-subscribed prefix code");
				width = maxLen;
				table.assign(std::size_t(1) << width, 0);
				uint32_t next[kMaxCodeLen + 2] = {0};
				uint32_t code = 0;
				for(int l = 1; l <= kMaxCodeLen; ++l) {
					code = (code + count[l - 1]) << 1;
					next[l] = code;
				}
				for(std::size_t s = 0; s < lens.size(); ++s) {
					const int l = lens[s];
					if(!l) continue;
					const uint32_t rev = reverseBits(next[l]++, l);
					const uint32_t entry = (static_cast<uint32_t>(s) << 4) | static_cast<uint32_t>(l);
					for(uint32_t i = rev; i < table.size(); i += (1u << l)) table[i] = entry;
				}
			}
			inline uint32_t decode(BitReader& br) const {
				if(width == 0) return single;
				const uint32_t e = table[br.peek(width)];
				br.skip(static_cast<int>(e & 15u));
				return e >> 4;
			}
		};

		void readCodeLengths(BitReader& br, int alphabet, std::vector<uint8_t>& lens) {
			lens.assign(alphabet, 0);
			if(br.read(1)) { // simple code: one or two symbols
				const int num = static_cast<int>(br.read(1)) + 1;
				const int firstBits = br.read(1) ? 8 : 1;
				const uint32_t s0 = br.read(firstBits);
				if(static_cast<int>(s0) >= alphabet) fail("simple code symbol outside the alphabet");
				lens[s0] = 1;
				if(num == 2) {
					const uint32_t s1 = br.read(8);
					if(static_cast<int>(s1) >= alphabet) fail("simple code symbol outside the alphabet");
					lens[s1] = 1;
				}
				return;
			}
			std::vector<uint8_t> clLens(kCodeLengthCodes, 0);
			const int numCl = 4 + static_cast<int>(br.read(4));
			if(numCl > kCodeLengthCodes) fail("too many code length codes");
			for(int i = 0; i < numCl; ++i) clLens[kCodeLengthOrder[i]] = static_cast<uint8_t>(br.read(3));
			PrefixCode cl;
			cl.build(clLens);
			int maxSymbol = alphabet;
			if(br.read(1)) {
				const int nbits = 2 + 2 * static_cast<int>(br.read(3));
				maxSymbol = 2 + static_cast<int>(br.read(nbits));
				if(maxSymbol > alphabet) fail("max_symbol larger than the alphabet");
			}
			int sym = 0, prev = 8;
			while(sym < alphabet) {
				if(maxSymbol-- == 0) break;
				const uint32_t c = cl.decode(br);
				if(c < 16) {
					lens[sym++] = static_cast<uint8_t>(c);
					if(c) prev = static_cast<int>(c);
				} else {
					static const int extra[3] = {2, 3, 7}, offset[3] = {3, 3, 11};
					const int rep = static_cast<int>(br.read(extra[c - 16])) + offset[c - 16];
					if(sym + rep > alphabet) fail("code length run past the alphabet");
					const uint8_t v = c == 16 ? static_cast<uint8_t>(prev) : 0;
					for(int i = 0; i < rep; ++i) lens[sym++] = v;
				}
				if(br.overrun) fail("truncated stream");
			}
		}

		// ------------------------------------------------------------------------------------------
		// LZ77 helpers
		// ------------------------------------------------------------------------------------------
		inline uint32_t prefixValue(uint32_t prefix, BitReader& br) {
			if(prefix < 4) return prefix + 1;
			const int extra = static_cast<int>(prefix - 2) >> 1;
			const uint32_t offset = (2u + (prefix & 1u)) << extra;
			return offset + br.read(extra) + 1;
		}

		// distance codes 1..120 address a neighbourhood of the current pixel: (dx, dy) packed as
		// (dy << 4) | (8 - dx)
		const uint8_t kCodeToPlane[120] = {
			0x18, 0x07, 0x17, 0x19, 0x28, 0x06, 0x27, 0x29, 0x16, 0x1a, 0x26, 0x2a, 0x38, 0x05, 0x37, 0x39, 0x15, 0x1b, 0x36, 0x3a,
			0x25, 0x2b, 0x48, 0x04, 0x47, 0x49, 0x14, 0x1c, 0x35, 0x3b, 0x46, 0x4a, 0x24, 0x2c, 0x58, 0x45, 0x4b, 0x34, 0x3c, 0x03,
			0x57, 0x59, 0x13, 0x1d, 0x56, 0x5a, 0x23, 0x2d, 0x44, 0x4c, 0x55, 0x5b, 0x33, 0x3d, 0x68, 0x02, 0x67, 0x69, 0x12, 0x1e,
			0x66, 0x6a, 0x22, 0x2e, 0x54, 0x5c, 0x43, 0x4d, 0x65, 0x6b, 0x32, 0x3e, 0x78, 0x01, 0x77, 0x79, 0x53, 0x5d, 0x11, 0x1f,
			0x64, 0x6c, 0x42, 0x4e, 0x76, 0x7a, 0x21, 0x2f, 0x75, 0x7b, 0x31, 0x3f, 0x63, 0x6d, 0x52, 0x5e, 0x00, 0x74, 0x7c, 0x41,
			0x4f, 0x10, 0x20, 0x62, 0x6e, 0x30, 0x73, 0x7d, 0x51, 0x5f, 0x40, 0x72, 0x7e, 0x61, 0x6f, 0x50, 0x71, 0x7f, 0x60, 0x70};

		inline std::size_t planeCodeToDistance(std::size_t xsize, uint32_t code) {
			if(code > 120) return code - 120;
			const int d = kCodeToPlane[code - 1];
			const long long dist = static_cast<long long>(d >> 4) * static_cast<long long>(xsize) + (8 - (d & 15));
			return dist >= 1 ? static_cast<std::size_t>(dist) : 1;
		}

		inline std::size_t subSize(std::size_t size, int bits) {
			return (size + (std::size_t(1) << bits) - 1) >> bits;
		}

		// ------------------------------------------------------------------------------------------
		// pixel arithmetic
		// ------------------------------------------------------------------------------------------
		inline uint32_t addPixels(uint32_t a, uint32_t b) {
			const uint32_t ag = (a & 0xff00ff00u) + (b & 0xff00ff00u);
			const uint32_t rb = (a & 0x00ff00ffu) + (b & 0x00ff00ffu);
			return (ag & 0xff00ff00u) | (rb & 0x00ff00ffu);
		}
		inline uint32_t subPixels(uint32_t a, uint32_t b) {
			const uint32_t ag = 0x00ff00ffu + (a & 0xff00ff00u) - (b & 0xff00ff00u);
			const uint32_t rb = 0xff00ff00u + (a & 0x00ff00ffu) - (b & 0x00ff00ffu);
			return (ag & 0xff00ff00u) | (rb & 0x00ff00ffu);
		}
		inline uint32_t avg2(uint32_t a, uint32_t b) {
			return (((a ^ b) & 0xfefefefeu) >> 1) + (a & b);
		}
		inline int ch(uint32_t p, int shift) {
			return static_cast<int>((p >> shift) & 0xff);
		}
		inline uint32_t select(uint32_t L, uint32_t T, uint32_t TL) {
			int pL = 0, pT = 0;
			for(int s = 0; s < 32; s += 8) {
				pL += std::abs(ch(T, s) - ch(TL, s)); // |predicted - L| with predicted = L + T - TL
				pT += std::abs(ch(L, s) - ch(TL, s));
			}
			return pL < pT ? L : T;
		}
		inline int clamp255(int v) {
			return v < 0 ? 0 : (v > 255 ? 255 : v);
		}
		inline uint32_t clampAddSubFull(uint32_t a, uint32_t b, uint32_t c) {
			uint32_t r = 0;
			for(int s = 0; s < 32; s += 8) r |= static_cast<uint32_t>(clamp255(ch(a, s) + ch(b, s) - ch(c, s))) << s;
			return r;
		}
		inline uint32_t clampAddSubHalf(uint32_t a, uint32_t b) {
			uint32_t r = 0;
			for(int s = 0; s < 32; s += 8) r |= static_cast<uint32_t>(clamp255(ch(a, s) + (ch(a, s) - ch(b, s)) / 2)) << s;
			return r;
		}
		inline uint32_t predict(int mode, uint32_t L, uint32_t T, uint32_t TR, uint32_t TL) {
			switch(mode) {
				case 0: return 0xff000000u;
				case 1: return L;
				case 2: return T;
				case 3: return TR;
				case 4: return TL;
				case 5: return avg2(avg2(L, TR), T);
				case 6: return avg2(L, TL);
				case 7: return avg2(L, T);
				case 8: return avg2(TL, T);
				case 9: return avg2(T, TR);
				case 10: return avg2(avg2(L, TL), avg2(T, TR));
				case 11: return select(L, T, TL);
				case 12: return clampAddSubFull(L, T, TL);
				case 13: return clampAddSubHalf(avg2(L, T), TL);
				default: return 0xff000000u; // modes 14, 15 behave like 0
			}
		}
		inline int8_t colorDelta(int8_t t, int8_t c) {
			return static_cast<int8_t>((static_cast<int>(t) * static_cast<int>(c)) >> 5);
		}

		// ------------------------------------------------------------------------------------------
		// decoder
		// ------------------------------------------------------------------------------------------
		struct Transform {
			int type = 0, bits = 0;
			std::size_t xsize = 0, ysize = 0; // size of the image this transform produces
			std::vector<uint32_t> data;       // sub-image or colour table
		};

		struct CodeGroup {
			PrefixCode c[5]; // green/length/cache, red, blue, alpha, distance
		};

		struct Decoder {
			BitReader br;
			Decoder(const uint8_t* d, std::size_t n) : br(d, n) {}

			std::vector<uint32_t> decodeImage(std::size_t xsize, std::size_t ysize, bool level0, std::vector<Transform>* transforms) {
				if(level0) {
					unsigned seen = 0;
					while(br.read(1)) {
						Transform t;
						t.type = static_cast<int>(br.read(2));
						if(seen & (1u << t.type)) fail("transform used twice");
						seen |= 1u << t.type;
						t.xsize = xsize;
						t.ysize = ysize;
						if(t.type == 0 || t.type == 1) {
							t.bits = static_cast<int>(br.read(3)) + 2;
							t.data = decodeImage(subSize(xsize, t.bits), subSize(ysize, t.bits), false, nullptr);
						} else if(t.type == 3) {
							const std::size_t colors = br.read(8) + 1;
							t.bits = colors > 16 ? 0 : (colors > 4 ? 1 : (colors > 2 ? 2 : 3));
							t.data = decodeImage(colors, 1, false, nullptr);
							for(std::size_t i = 1; i < colors; ++i) t.data[i] = addPixels(t.data[i], t.data[i - 1]);
							// indices past the table decode to transparent black
							t.data.resize(std::size_t(1) << (8 >> t.bits), 0u);
							xsize = subSize(xsize, t.bits);
						}
						transforms->push_back(std::move(t));
						if(br.overrun) fail("truncated stream");
					}
				}
				int cacheBits = 0;
				if(br.read(1)) {
					cacheBits = static_cast<int>(br.read(4));
					if(cacheBits < 1 || cacheBits > 11) fail("invalid colour cache size");
				}
				std::vector<uint32_t> meta;
				int metaBits = 0;
				std::size_t metaW = 0, groups = 1;
				if(level0 && br.read(1)) {
					metaBits = static_cast<int>(br.read(3)) + 2;
					metaW = subSize(xsize, metaBits);
					meta = decodeImage(metaW, subSize(ysize, metaBits), false, nullptr);
					uint32_t mx = 0;
					for(uint32_t& m : meta) {
						m = (m >> 8) & 0xffffu;
						mx = std::max(mx, m);
					}
					groups = static_cast<std::size_t>(mx) + 1;
				}
				std::vector<CodeGroup> codes(groups);
				std::vector<uint8_t> lens;
				const int alphabets[5] = {256 + 24 + (cacheBits ? (1 << cacheBits) : 0), 256, 256, 256, 40};
				for(CodeGroup& g : codes) {
					for(int k = 0; k < 5; ++k) {
						readCodeLengths(br, alphabets[k], lens);
						g.c[k].build(lens);
					}
					if(br.overrun) fail("truncated stream");
				}

				const std::size_t total = xsize * ysize;
				std::vector<uint32_t> out(total);
				std::vector<uint32_t> cache(cacheBits ? (std::size_t(1) << cacheBits) : 0);
				const int cacheShift = 32 - cacheBits;
				auto remember = [&](uint32_t argb) {
					if(cacheBits) cache[(0x1e35a7bdu * argb) >> cacheShift] = argb;
				};
				std::size_t pos = 0, x = 0, y = 0;
				const CodeGroup* g = &codes[0];
				const std::size_t metaMask = metaBits ? ((std::size_t(1) << metaBits) - 1) : ~std::size_t(0);
				bool pick = true;
				while(pos < total) {
					if(metaBits && pick) g = &codes[meta[(y >> metaBits) * metaW + (x >> metaBits)]];
					pick = false;
					const uint32_t s = g->c[0].decode(br);
					if(s < 256) {
						const uint32_t r = g->c[1].decode(br);
						const uint32_t b = g->c[2].decode(br);
						const uint32_t a = g->c[3].decode(br);
						const uint32_t px = (a << 24) | (r << 16) | (s << 8) | b;
						out[pos++] = px;
						remember(px);
						if(++x == xsize) {
							x = 0;
							++y;
							pick = true;
						} else if((x & metaMask) == 0) {
							pick = true;
						}
					} else if(s < 256 + 24) {
						const std::size_t length = prefixValue(s - 256, br);
						const uint32_t dsym = g->c[4].decode(br);
						const std::size_t dist = planeCodeToDistance(xsize, prefixValue(dsym, br));
						if(dist > pos || length > total - pos) fail("backward reference outside the image");
						for(std::size_t i = 0; i < length; ++i, ++pos) {
							out[pos] = out[pos - dist];
							remember(out[pos]);
						}
						x += length;
						while(x >= xsize) {
							x -= xsize;
							++y;
						}
						pick = true;
					} else {
						const std::size_t idx = s - (256 + 24);
						if(idx >= cache.size()) fail("colour cache index out of range");
						const uint32_t px = cache[idx];
						out[pos++] = px;
						remember(px);
						if(++x == xsize) {
							x = 0;
							++y;
							pick = true;
						} else if((x & metaMask) == 0) {
							pick = true;
						}
					}
					if(br.overrun) fail("truncated stream");
				}
				return out;
			}
		};

		void inversePredictor(const Transform& t, std::vector<uint32_t>& px) {
			const std::size_t w = t.xsize, h = t.ysize, tw = subSize(w, t.bits);
			if(w == 0 || h == 0) return;
			px[0] = addPixels(px[0], 0xff000000u);
			for(std::size_t x = 1; x < w; ++x) px[x] = addPixels(px[x], px[x - 1]);
			for(std::size_t y = 1; y < h; ++y) {
				uint32_t* row = &px[y * w];
				const uint32_t* up = row - w;
				row[0] = addPixels(row[0], up[0]);
				for(std::size_t x = 1; x < w; ++x) {
					const int mode = static_cast<int>((t.data[(y >> t.bits) * tw + (x >> t.bits)] >> 8) & 0xf);
					// up[x + 1] of the last column is the first pixel of the current row (memory order)
					row[x] = addPixels(row[x], predict(mode, row[x - 1], up[x], up[x + 1], up[x - 1]));
				}
			}
		}

		void inverseColor(const Transform& t, std::vector<uint32_t>& px) {
			const std::size_t w = t.xsize, h = t.ysize, tw = subSize(w, t.bits);
			for(std::size_t y = 0; y < h; ++y) {
				for(std::size_t x = 0; x < w; ++x) {
					const uint32_t e = t.data[(y >> t.bits) * tw + (x >> t.bits)];
					const int8_t g2r = static_cast<int8_t>(e & 0xff), g2b = static_cast<int8_t>((e >> 8) & 0xff), r2b = static_cast<int8_t>((e >> 16) & 0xff);
					const uint32_t p = px[y * w + x];
					const int8_t green = static_cast<int8_t>((p >> 8) & 0xff);
					int red = static_cast<int>((p >> 16) & 0xff), blue = static_cast<int>(p & 0xff);
					red = (red + colorDelta(g2r, green)) & 0xff;
					blue = (blue + colorDelta(g2b, green)) & 0xff;
					blue = (blue + colorDelta(r2b, static_cast<int8_t>(red))) & 0xff;
					px[y * w + x] = (p & 0xff00ff00u) | (static_cast<uint32_t>(red) << 16) | static_cast<uint32_t>(blue);
				}
			}
		}

		void inverseSubtractGreen(std::vector<uint32_t>& px) {
			for(uint32_t& p : px) {
				const uint32_t g = (p >> 8) & 0xff;
				const uint32_t rb = ((p & 0x00ff00ffu) + ((g << 16) | g)) & 0x00ff00ffu;
				p = (p & 0xff00ff00u) | rb;
			}
		}

		void inverseColorIndexing(const Transform& t, std::vector<uint32_t>& px) {
			const std::size_t w = t.xsize, h = t.ysize, sw = subSize(w, t.bits);
			const int bpp = 8 >> t.bits;
			const uint32_t mask = (1u << bpp) - 1;
			const std::size_t perPixel = (std::size_t(1) << t.bits) - 1;
			std::vector<uint32_t> out(w * h);
			for(std::size_t y = 0; y < h; ++y) {
				for(std::size_t x = 0; x < w; ++x) {
					const uint32_t packed = (px[y * sw + (x >> t.bits)] >> 8) & 0xff;
					const uint32_t idx = (packed >> ((x & perPixel) * bpp)) & mask;
					out[y * w + x] = t.data[idx]; // the table was padded to 2^bpp entries (bits > 0) or 256
				}
			}
			px.swap(out);
		}

		// ------------------------------------------------------------------------------------------
		// encoder helpers
		// ------------------------------------------------------------------------------------------
		// Code lengths of a Huffman code for `hist`, no longer than maxLen.  Counts below a rising
		// floor are lifted to it until the tree is shallow enough (the tree stays complete).
		std::vector<uint8_t> huffmanLengths(const std::vector<uint32_t>& hist, int maxLen) {
			const std::size_t n = hist.size();
			std::vector<uint8_t> lens(n, 0);
			std::vector<std::size_t> used;
			for(std::size_t s = 0; s < n; ++s)
				if(hist[s]) used.push_back(s);
			if(used.empty()) return lens;
			if(used.size() == 1) {
				lens[used[0]] = 1;
				return lens;
			}
			for(uint64_t floor = 1;; floor *= 2) {
				struct Node {
					uint64_t w;
					int left, right;
				};
				std::vector<Node> nodes;
				using Item = std::pair<uint64_t, int>;
				std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
				for(std::size_t s : used) {
					nodes.push_back({std::max<uint64_t>(hist[s], floor), -1, -1});
					heap.push({nodes.back().w, static_cast<int>(nodes.size()) - 1});
				}
				while(heap.size() > 1) {
					const Item a = heap.top();
					heap.pop();
					const Item b = heap.top();
					heap.pop();
					nodes.push_back({a.first + b.first, a.second, b.second});
					heap.push({a.first + b.first, static_cast<int>(nodes.size()) - 1});
				}
				// depth of every leaf
				std::vector<int> depth(nodes.size(), 0);
				int deepest = 0;
				for(int i = static_cast<int>(nodes.size()) - 1; i >= 0; --i) {
					if(nodes[i].left >= 0) {
						depth[nodes[i].left] = depth[i] + 1;
						depth[nodes[i].right] = depth[i] + 1;
					} else {
						deepest = std::max(deepest, depth[i]);
					}
				}
				if(deepest <= maxLen) {
					for(std::size_t k = 0; k < used.size(); ++k) lens[used[k]] = static_cast<uint8_t>(depth[k]);
					return lens;
				}
			}
		}

		struct EncCode {
			std::vector<uint8_t> lens;
			std::vector<uint32_t> bits; // already bit-reversed for the LSB-first writer
			void assign(const std::vector<uint8_t>& l) {
				lens = l;
				bits.assign(l.size(), 0);
				int count[kMaxCodeLen + 1] = {0};
				for(uint8_t v : l) ++count[v];
				count[0] = 0;
				uint32_t next[kMaxCodeLen + 2] = {0}, code = 0;
				for(int k = 1; k <= kMaxCodeLen; ++k) {
					code = (code + count[k - 1]) << 1;
					next[k] = code;
				}
				for(std::size_t s = 0; s < l.size(); ++s)
					if(l[s]) bits[s] = reverseBits(next[l[s]]++, l[s]);
			}
			inline void put(BitWriter& bw, uint32_t sym) const {
				bw.put(bits[sym], lens[sym]);
			}
		};

		// Builds the code for `hist` and writes its description; symbols of a one-symbol code cost no bits.
		void writeCode(BitWriter& bw, const std::vector<uint32_t>& hist, EncCode& code) {
			std::vector<std::size_t> used;
			for(std::size_t s = 0; s < hist.size(); ++s)
				if(hist[s]) used.push_back(s);
			if(used.size() <= 2 && (used.empty() || used.back() < 256)) {
				// simple code
				std::vector<uint8_t> lens(hist.size(), 0);
				const std::size_t s0 = used.empty() ? 0 : used[0];
				bw.put(1, 1);
				bw.put(used.size() == 2 ? 1 : 0, 1);
				if(s0 < 2) {
					bw.put(0, 1);
					bw.put(static_cast<uint32_t>(s0), 1);
				} else {
					bw.put(1, 1);
					bw.put(static_cast<uint32_t>(s0), 8);
				}
				if(used.size() == 2) {
					bw.put(static_cast<uint32_t>(used[1]), 8);
					lens[s0] = lens[used[1]] = 1;
				}
				code.assign(lens); // one symbol: length 0, nothing is written per pixel
				return;
			}
			std::vector<uint8_t> lens = huffmanLengths(hist, kMaxCodeLen);
			if(used.size() == 1) {
				// one symbol that does not fit a simple code: give it a partner so the code is complete
				lens[used[0]] = 1;
				lens[used[0] == 0 ? 1 : 0] = 1;
			}
			code.assign(lens);
			// the lengths themselves, coded with literal length symbols 0..15 only
			std::vector<uint32_t> clHist(kCodeLengthCodes, 0);
			for(uint8_t l : lens) ++clHist[l];
			int distinct = 0;
			for(uint32_t c : clHist) distinct += c != 0;
			if(distinct == 1) clHist[clHist[0] ? 1 : 0] = 1; // keep the code-length code complete
			EncCode cl;
			cl.assign(huffmanLengths(clHist, 7));
			bw.put(0, 1);                    // normal code
			bw.put(kCodeLengthCodes - 4, 4); // all 19 code length code lengths follow
			for(int i = 0; i < kCodeLengthCodes; ++i) bw.put(cl.lens[kCodeLengthOrder[i]], 3);
			bw.put(0, 1); // max_symbol = alphabet size
			for(uint8_t l : lens) cl.put(bw, l);
		}

		// An entropy-coded image without LZ77 or colour cache: five codes, then one literal per pixel.
		void writeLiteralImage(BitWriter& bw, const std::vector<uint32_t>& px) {
			std::vector<uint32_t> hg(256 + 24, 0), hr(256, 0), hb(256, 0), ha(256, 0), hd(40, 0);
			for(uint32_t p : px) {
				++hg[(p >> 8) & 0xff];
				++hr[(p >> 16) & 0xff];
				++hb[p & 0xff];
				++ha[p >> 24];
			}
			EncCode cg, cr, cb, ca, cd;
			writeCode(bw, hg, cg);
			writeCode(bw, hr, cr);
			writeCode(bw, hb, cb);
			writeCode(bw, ha, ca);
			writeCode(bw, hd, cd);
			for(uint32_t p : px) {
				cg.put(bw, (p >> 8) & 0xff);
				cr.put(bw, (p >> 16) & 0xff);
				cb.put(bw, p & 0xff);
				ca.put(bw, p >> 24);
			}
		}

		inline int residualCost(uint32_t r) {
			int c = 0;
			for(int s = 0; s < 32; s += 8) {
				const int v = ch(r, s);
				c += v < 128 ? v : 256 - v;
			}
			return c;
		}
	}

	bool IsWebP(const uint8_t* d, std::size_t n) {
		return n >= 12 && std::memcmp(d, "RIFF", 4) == 0 && std::memcmp(d + 8, "WEBP", 4) == 0;
	}

	Decoded Decode(const uint8_t* data, std::size_t size) {
		if(!IsWebP(data, size)) fail("not a RIFF/WEBP file");
		// walk the chunks for the VP8L payload (simple files have it first; extended ones after VP8X)
		std::size_t off = 12;
		const uint8_t* payload = nullptr;
		std::size_t payloadSize = 0;
		while(off + 8 <= size) {
			const uint32_t csize = rd32(data + off + 4);
			if(std::memcmp(data + off, "VP8L", 4) == 0) {
				payload = data + off + 8;
				payloadSize = std::min<std::size_t>(csize, size - off - 8);
				break;
			}
			if(std::memcmp(data + off, "VP8 ", 4) == 0) fail("lossy (VP8) data is not supported by this build");
			if(std::memcmp(data + off, "ANIM", 4) == 0 || std::memcmp(data + off, "ANMF", 4) == 0) fail("Animated WebP images are not supported!");
			off += 8 + static_cast<std::size_t>(csize) + (csize & 1u);
		}
		if(!payload || payloadSize < 5) fail("no VP8L chunk found");
		if(payload[0] != 0x2f) fail("bad VP8L signature");
		Decoder dec(payload + 1, payloadSize - 1);
		Decoded out;
		out.w = dec.br.read(14) + 1;
		out.h = dec.br.read(14) + 1;
		const bool alpha = dec.br.read(1) != 0;
		if(dec.br.read(3) != 0) fail("unknown VP8L version");
		std::vector<Transform> transforms;
		std::vector<uint32_t> px = dec.decodeImage(out.w, out.h, true, &transforms);
		for(std::size_t k = transforms.size(); k-- > 0;) {
			const Transform& t = transforms[k];
			switch(t.type) {
				case 0: inversePredictor(t, px); break;
				case 1: inverseColor(t, px); break;
				case 2: inverseSubtractGreen(px); break;
				default: inverseColorIndexing(t, px); break;
			}
		}
		out.channels = alpha ? 4 : 3;
		out.pixels.resize(static_cast<std::size_t>(out.w) * out.h * out.channels);
		uint8_t* o = out.pixels.data();
		for(uint32_t p : px) {
			*o++ = static_cast<uint8_t>(p >> 16);
			*o++ = static_cast<uint8_t>(p >> 8);
			*o++ = static_cast<uint8_t>(p);
			if(alpha) *o++ = static_cast<uint8_t>(p >> 24);
		}
		return out;
	}

	std::vector<uint8_t> Encode(const uint8_t* pixels, unsigned w, unsigned h, int channels) {
		if(w < 1 || h < 1 || w > 16384 || h > 16384) fail("image size outside 1..16384");
		if(channels != 3 && channels != 4) fail("only RGB and RGBA can be encoded");
		const std::size_t total = static_cast<std::size_t>(w) * h;
		std::vector<uint32_t> px(total);
		bool alphaUsed = false;
		for(std::size_t i = 0; i < total; ++i) {
			const uint8_t* p = pixels + i * channels;
			const uint32_t a = channels == 4 ? p[3] : 0xffu;
			alphaUsed = alphaUsed || a != 0xffu;
			px[i] = (a << 24) | (static_cast<uint32_t>(p[0]) << 16) | (static_cast<uint32_t>(p[1]) << 8) | p[2];
		}
		BitWriter bw;
		bw.put(w - 1, 14);
		bw.put(h - 1, 14);
		// As libwebp does, the alpha hint is set only when some pixel is not opaque; the reference's
		// decoder maps it to Image::Layout (libs/image/src/WebP.cpp:52), so fully opaque RGBA faces
		// come back as RGB there, and here.
		bw.put(alphaUsed ? 1 : 0, 1);
		bw.put(0, 3);

		// transform 1: subtract green
		bw.put(1, 1);
		bw.put(2, 2);
		for(uint32_t& p : px) {
			const uint32_t g = (p >> 8) & 0xff;
			const uint32_t rb = ((p & 0x00ff00ffu) + 0x01000100u - ((g << 16) | g)) & 0x00ff00ffu;
			p = (p & 0xff00ff00u) | rb;
		}

		// transform 2: spatial prediction, one mode per 32 x 32 tile, chosen by the smallest sum of
		// absolute residuals
		const int bits = 5;
		const std::size_t tw = subSize(w, bits), th = subSize(h, bits);
		std::vector<uint32_t> modes(tw * th, 0xff000000u);
		std::vector<uint32_t> res(total);
		auto neighbours = [&](std::size_t x, std::size_t y, uint32_t& L, uint32_t& T, uint32_t& TR, uint32_t& TL) {
			const uint32_t* row = &px[y * w];
			L = row[x - 1];
			T = row[x - w];
			TL = row[x - w - 1];
			TR = (x + 1 < w) ? row[x + 1 - w] : row[0];
		};
		for(std::size_t ty = 0; ty < th; ++ty) {
			for(std::size_t tx = 0; tx < tw; ++tx) {
				const std::size_t x0 = std::max<std::size_t>(tx << bits, 1), x1 = std::min<std::size_t>((tx + 1) << bits, w);
				const std::size_t y0 = std::max<std::size_t>(ty << bits, 1), y1 = std::min<std::size_t>((ty + 1) << bits, h);
				long long best = -1;
				int bestMode = 1;
				for(int mode = 0; mode < 14; ++mode) {
					long long cost = 0;
					for(std::size_t y = y0; y < y1; ++y)
						for(std::size_t x = x0; x < x1; ++x) {
							uint32_t L, T, TR, TL;
							neighbours(x, y, L, T, TR, TL);
							cost += residualCost(subPixels(px[y * w + x], predict(mode, L, T, TR, TL)));
						}
					if(best < 0 || cost < best) {
						best = cost;
						bestMode = mode;
					}
				}
				modes[ty * tw + tx] = 0xff000000u | (static_cast<uint32_t>(bestMode) << 8);
			}
		}
		res[0] = subPixels(px[0], 0xff000000u);
		for(std::size_t x = 1; x < w; ++x) res[x] = subPixels(px[x], px[x - 1]);
		for(std::size_t y = 1; y < h; ++y) {
			res[y * w] = subPixels(px[y * w], px[(y - 1) * w]);
			for(std::size_t x = 1; x < w; ++x) {
				uint32_t L, T, TR, TL;
				neighbours(x, y, L, T, TR, TL);
				const int mode = static_cast<int>((modes[(y >> bits) * tw + (x >> bits)] >> 8) & 0xf);
				res[y * w + x] = subPixels(px[y * w + x], predict(mode, L, T, TR, TL));
			}
		}
		bw.put(1, 1);
		bw.put(0, 2);
		bw.put(bits - 2, 3);
		bw.put(0, 1); // sub-image: no colour cache (and no meta prefix bit below level 0)
		writeLiteralImage(bw, modes);

		bw.put(0, 1); // no more transforms
		bw.put(0, 1); // no colour cache
		bw.put(0, 1); // one prefix code group
		writeLiteralImage(bw, res);
		bw.flush();

		std::vector<uint8_t> file;
		const uint32_t chunk = static_cast<uint32_t>(bw.out.size()) + 1;
		const uint32_t padded = chunk + (chunk & 1u);
		auto put32 = [&file](uint32_t v) {
			for(int i = 0; i < 4; ++i) file.push_back(static_cast<uint8_t>(v >> (8 * i)));
		};
		file.insert(file.end(), {'R', 'I', 'F', 'F'});
		put32(4 + 8 + padded);
		file.insert(file.end(), {'W', 'E', 'B', 'P', 'V', 'P', '8', 'L'});
		put32(chunk);
		file.push_back(0x2f);
		file.insert(file.end(), bw.out.begin(), bw.out.end());
		if(chunk & 1u) file.push_back(0);
		return file;
	}
}
