// datasetGen N p file -- Erdos-Renyi edge-list writer, format-compatible with the reference's
// src/datasetGenerator.cpp (header "N\tE", then "name\tname\tweight" per undirected edge, upper triangle only,
// std::default_random_engine(10000), 12-character random node names from libc rand()).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <random>
#include <set>
#include <string>
#include <vector>

static std::vector<std::string> generateRandomName(const int n) {     // datasetGenerator.cpp:163-192
	const char alphanum[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz";
	std::set<std::string> names;
	const int slen = 12;
	char buf[slen + 1];
	buf[slen] = 0;
	while ((int)names.size() < n) {
		for (int i = 0; i < slen; i++) buf[i] = alphanum[rand() % (sizeof(alphanum) - 1)];
		names.emplace(buf);
	}
	return std::vector<std::string>(names.begin(), names.end());
}

int main(int argc, const char ** argv) {
	if (argc < 4) { std::cout << "usage: ./datasetGenerator numberOfNodes edgeProbability outputFilename" << std::endl; return -1; }
	const uint32_t nNodes = atoi(argv[1]);
	const float probDensity = atof(argv[2]);
	const std::string netFileName(argv[3]);
	std::cout << "*** Erdos graph generator ***" << std::endl;
	std::cout << "number of nodes: " << nNodes << " - edge probability: " << probDensity << " - output filename: " << netFileName << std::endl;
	std::default_random_engine eng(10000);                             // :38,41
	std::uniform_real_distribution<> randR(0.0, 1.0);
	std::ofstream netFile(netFileName.c_str(), std::ios::out);
	if (!netFile.is_open()) { std::cerr << "error opening output file" << std::endl; return -1; }
	std::vector<std::string> nodeNames = generateRandomName(nNodes);
	std::vector<std::vector<uint32_t>> edges(nNodes);
	uint64_t nEdges = 0;
	for (uint32_t i = 0; i + 1 < nNodes; i++)                          // :74-98
		for (uint32_t j = i + 1; j < nNodes; j++)
			if (randR(eng) < probDensity) { edges[i].push_back(j); nEdges++; }
	std::cout << "nEdges: " << nEdges << std::endl;
	netFile << nNodes << "\t" << nEdges << std::endl;                  // :147
	for (uint32_t i = 0; i < nNodes; i++)
		for (uint32_t j : edges[i])
			netFile << nodeNames[i] << "\t" << nodeNames[j] << "\t" << randR(eng) << std::endl;   // :153
	netFile.close();
	return 0;
}
