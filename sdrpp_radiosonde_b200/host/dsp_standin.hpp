/*
 * dsp_standin.hpp — minimal stand-in for the parts of SDR++ core the decoder block touches.
 *
 * SDR++ (AlexandreRouma/SDRPlusPlus) is not vendored in the reference tree and not available
 * offline, so the block in gpu_decoder.hpp is compiled against this API-compatible subset
 * (same names, same call protocol) for tests; define SONDE_B200_USE_SDRPP to compile against
 * the real <dsp/block.h> / <dsp/stream.h> instead.
 *
 * Mirrored surface (as used by src/decode/decoder.hpp:23-59,117 and src/main.cpp:57-68):
 *   dsp::complex_t {re, im}
 *   dsp::stream<T>: writeBuf, readBuf, swap(n), read(), flush(), stopWriter(), stopReader(), clear*Stop()
 *   dsp::block    : registerInput/unregisterInput, start(), stop(), virtual int run(), _block_init
 */
#pragma once
#ifdef SONDE_B200_USE_SDRPP
#include <dsp/block.h>
#include <dsp/stream.h>
#include <dsp/types.h>
#else
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace dsp {

struct complex_t {
	float re, im;
};

constexpr int STREAM_BUFFER_SIZE = 1000000;

class untyped_stream {
public:
	virtual ~untyped_stream() {}
	virtual void stopWriter() = 0;
	virtual void clearWriteStop() = 0;
	virtual void stopReader() = 0;
	virtual void clearReadStop() = 0;
};

/* double-buffer hand-off between a producer and a consumer thread */
template <class T>
class stream : public untyped_stream {
public:
	stream() : writeBuf((T *)malloc(STREAM_BUFFER_SIZE * sizeof(T))), readBuf((T *)malloc(STREAM_BUFFER_SIZE * sizeof(T))) {}
	~stream() { free(writeBuf); free(readBuf); }

	/* producer: publish `size` items of writeBuf; blocks until the consumer has flushed the previous buffer */
	bool swap(int size)
	{
		{
			std::unique_lock<std::mutex> lck(swapMtx);
			swapCV.wait(lck, [this] { return canSwap || writerStop; });
			if (writerStop) return false;
			dataSize = size;
			canSwap = false;
			std::swap(writeBuf, readBuf);
		}
		{
			std::lock_guard<std::mutex> lck(rdyMtx);
			dataReady = true;
		}
		rdyCV.notify_all();
		return true;
	}

	/* consumer: wait for data; returns the item count or -1 when stopped */
	int read()
	{
		std::unique_lock<std::mutex> lck(rdyMtx);
		rdyCV.wait(lck, [this] { return dataReady || readerStop; });
		return readerStop ? -1 : dataSize;
	}

	void flush()
	{
		{
			std::lock_guard<std::mutex> lck(rdyMtx);
			dataReady = false;
		}
		{
			std::lock_guard<std::mutex> lck(swapMtx);
			canSwap = true;
		}
		swapCV.notify_all();
	}

	void stopWriter() override { { std::lock_guard<std::mutex> l(swapMtx); writerStop = true; } swapCV.notify_all(); }
	void clearWriteStop() override { writerStop = false; }
	void stopReader() override { { std::lock_guard<std::mutex> l(rdyMtx); readerStop = true; } rdyCV.notify_all(); }
	void clearReadStop() override { readerStop = false; }

	T *writeBuf;
	T *readBuf;

private:
	std::mutex swapMtx, rdyMtx;
	std::condition_variable swapCV, rdyCV;
	bool canSwap = true, dataReady = false, readerStop = false, writerStop = false;
	int dataSize = 0;
};

class block {
public:
	virtual ~block() {}
	virtual int run() = 0;

	void start()
	{
		if (running) return;
		running = true;
		for (auto *in : inputs) in->clearReadStop();
		worker = std::thread([this] { while (run() >= 0) {} });
	}

	void stop()
	{
		if (!running) return;
		for (auto *in : inputs) in->stopReader();
		if (worker.joinable()) worker.join();
		running = false;
	}

protected:
	void registerInput(untyped_stream *s) { inputs.push_back(s); }
	void unregisterInput(untyped_stream *s)
	{
		for (size_t i = 0; i < inputs.size(); i++)
			if (inputs[i] == s) { inputs.erase(inputs.begin() + i); break; }
	}
	bool _block_init = false;
	bool running = false;

private:
	std::vector<untyped_stream *> inputs;
	std::thread worker;
};

}  // namespace dsp
#endif
